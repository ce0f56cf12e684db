// Microbenchmark: FFMA2 issue rate vs warps per scheduler and independent chains per warp.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
#define FMA2(d,a,b,c) asm volatile("fma.rn.f32x2 %0, %1, %2, %3;":"=l"(d):"l"(a),"l"(b),"l"(c))
__device__ __forceinline__ u64 splat(float v){u64 d; asm("mov.b64 %0, {%1,%1};":"=l"(d):"f"(v)); return d;}
template<int CH> __global__ void k(float* out, int iters, float seed, long long* cyc){
  u64 acc[CH]; for(int i=0;i<CH;i++) acc[i]=splat(seed+i+threadIdx.x*1e-3f);
  u64 m=splat(0.9999f+threadIdx.x*1e-7f), c=splat(1e-9f+threadIdx.x*1e-12f);
  long long t0=clock64();
  for(int it=0;it<iters;++it){ _Pragma("unroll") for(int i=0;i<CH;i++) FMA2(acc[i],acc[i],m,c); }
  long long t1=clock64();
  float s=0; for(int i=0;i<CH;i++) s+=__uint_as_float((unsigned)acc[i]);
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
  if(threadIdx.x==0&&blockIdx.x==0) *cyc=t1-t0;
}
int main(){
  float* out; long long* cyc; cudaMalloc(&out,148*1024*4); cudaMallocManaged(&cyc,8);
  const int iters=20000;
  printf("cycles per FFMA2 per scheduler (2.0 = pipe peak)\n%-10s", "warps/sch");
  for(int ch:{4,8,16,32}) printf("  chains=%-3d",ch); printf("\n");
  for(int wps:{1,2,4,8}){
    printf("%-10d", wps);
    const int threads=wps*4*32;
    #define RUN(CH) for(int r=0;r<2;r++){ k<CH><<<148,threads>>>(out,iters,1.f,cyc); cudaDeviceSynchronize(); } printf("  %10.2f", (double)*cyc/iters/CH/wps);
    RUN(4) RUN(8) RUN(16) RUN(32)
    printf("\n");
  }
  printf("err=%s\n",cudaGetErrorString(cudaGetLastError()));
}
