#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <string.h>
int main(){
  const float c = 0.49999997f;
  uint64_t bad=0, bad_old=0;
  for(uint64_t u=0; u<(1ull<<32); ++u){ uint32_t b=(uint32_t)u; float x; memcpy(&x,&b,4); if(isnan(x)) continue;
    float r = roundf(x); float q = truncf(x + copysignf(c,x)); float o = truncf(x + copysignf(0.5f,x));
    if(!(r==q) || (signbit(r)!=signbit(q))) { if(bad<5) printf("new mismatch %a: %a vs %a\n",x,r,q); ++bad; }
    if(!(r==o)) ++bad_old; }
  printf("c=%a new mismatches %llu, old mismatches %llu\n", c, (unsigned long long)bad, (unsigned long long)bad_old);
  return 0; }
