// oracle/ref_shim: TEST INFRASTRUCTURE ONLY.
// Empty stand-in for the reference's flan/Function.h. The reference's
// WindowFunctions.cpp:4 includes it without using anything from it; the real
// header drags in Graph/Color/bitmap code that is not on the phase-vocoder path.
#pragma once
