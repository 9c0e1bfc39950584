// flan/defines.h of the B200 build: the scalar vocabulary of the reference (src/flan/defines.h:10-62),
// restated for the phase-vocoder path. Same names, widths and constants so user code compiles unchanged.
#pragma once

#include <atomic>
#include <cmath>
#include <cstdint>
#include <iostream>

namespace flan {

using Index = int;
using Second = float;
using Channel = int32_t;
using Frame = int32_t;
using Bin = int32_t;
using fFrame = float;
using fBin = float;
using Sample = float;
using Frequency = float;
using Magnitude = float;
using FrameRate = float;
using Radian = float;

struct MF { Magnitude m; Frequency f; };      // defines.h:29-33: 8 bytes, the element of PVBuffer
struct TF { Second t; Frequency f; };

const Radian pi = std::acos( -1.0f );          // defines.h:44
const Radian pi2 = pi * 2.0f;                  // defines.h:45

}

// Cooperative cancellation, defines.h:52-62. The B200 engine polls the flag between kernel launches.
#define flan_CANCELLABLE
static std::atomic<bool> default_canceller( false );
#define flan_CANCEL_POINT( T ) { if( canceller ) return T; }
#define flan_CANCEL_ARG std::atomic<bool> & = default_canceller
#define flan_CANCEL_ARG_CPP std::atomic<bool> & canceller
