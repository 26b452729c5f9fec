#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <math.h>
int main(){
  const float y = 1.0f/6.0f;
  long bad=0, badn=0; 
  #pragma omp parallel for reduction(+:bad,badn) schedule(static)
  for (int64_t b=0;b<(1LL<<32);++b){
    uint32_t u=(uint32_t)b; float x; memcpy(&x,&u,4);
    if (isnan(x)||isinf(x)) continue;
    float ref = x/6.0f;
    float q = x*y; float r = fmaf(-6.0f,q,x); float q2 = fmaf(r,y,q);
    if (q2!=ref) { bad++; if (fabsf(x) > 1e-30f && fabsf(x) < 1e30f) badn++; }
  }
  printf("mismatches %ld, in normal range %ld\n", bad, badn);
  return 0;
}
