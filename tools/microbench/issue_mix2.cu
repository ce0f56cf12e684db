// Microbenchmark: 8 FFMA2 per iteration mixed with broadcast LDS.128 / LDS.64 / SHF, independent of the FMAs.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
#define FMA2(d,a,b,c) asm volatile("fma.rn.f32x2 %0, %1, %2, %3;":"=l"(d):"l"(a),"l"(b),"l"(c))
__device__ __forceinline__ u64 splat(float v){u64 d; asm("mov.b64 %0, {%1,%1};":"=l"(d):"f"(v)); return d;}
template<int MODE> __global__ void __launch_bounds__(512) k(float* out, int iters, float seed, long long* cyc){
  __shared__ __align__(16) float sm[4096];
  for(int i=threadIdx.x;i<4096;i+=512) sm[i]=i; __syncthreads();
  const unsigned sb=(unsigned)__cvta_generic_to_shared(sm);
  u64 acc[8]; for(int i=0;i<8;i++) acc[i]=splat(seed+i+threadIdx.x*1e-3f);
  u64 m=splat(0.9999f+threadIdx.x*1e-7f), c=splat(1e-9f);
  u64 sink=0; unsigned sh=threadIdx.x;
  long long t0=clock64();
  for(int it=0;it<iters;++it){
    const unsigned a=sb+((it&63)<<6);
    #pragma unroll
    for(int i=0;i<8;i++){
      FMA2(acc[i],acc[i],m,c);
      if (MODE==1 && (i&3)==0){ u64 x,y; asm volatile("ld.shared.v2.b64 {%0,%1}, [%2];":"=l"(x),"=l"(y):"r"(a+16*(i>>2))); sink^=x^y; }
      if (MODE==2 && (i&1)==0){ u64 x; asm volatile("ld.shared.b64 %0, [%1];":"=l"(x):"r"(a+8*(i>>1))); sink^=x; }
      if (MODE==3 && (i&3)==0){ asm volatile("shf.l.wrap.b32 %0, %1, %0, 1;":"+r"(sh):"r"((unsigned)acc[i])); }
      if (MODE==4 && (i&3)==0){ u64 x,y; asm volatile("ld.shared.v2.b64 {%0,%1}, [%2];":"=l"(x),"=l"(y):"r"(a+16*(i>>2))); sink^=x^y; asm volatile("shf.l.wrap.b32 %0, %1, %0, 1;":"+r"(sh):"r"((unsigned)acc[i])); }
      if (MODE==5 && (i&3)==0){ float x; asm volatile("ld.shared.f32 %0, [%1];":"=f"(x):"r"(a+4*(i>>2))); sink^=__float_as_uint(x); }
    }
  }
  long long t1=clock64();
  float s=(float)sink+sh; for(int i=0;i<8;i++) s+=__uint_as_float((unsigned)acc[i]);
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
  if(threadIdx.x==0&&blockIdx.x==0) *cyc=t1-t0;
}
int main(){
  float* out; long long* cyc; cudaMalloc(&out,148*512*4); cudaMallocManaged(&cyc,8);
  const int iters=20000; const char* names[]={"8 FFMA2","8 FFMA2 + 2 LDS.128","8 FFMA2 + 4 LDS.64","8 FFMA2 + 2 SHF","8 FFMA2 + 2 LDS.128 + 2 SHF","8 FFMA2 + 2 LDS.32"};
  for(int mode=0;mode<6;mode++){ for(int r=0;r<2;r++){ switch(mode){case 0:k<0><<<148,512>>>(out,iters,1.f,cyc);break;case 1:k<1><<<148,512>>>(out,iters,1.f,cyc);break;case 2:k<2><<<148,512>>>(out,iters,1.f,cyc);break;case 3:k<3><<<148,512>>>(out,iters,1.f,cyc);break;case 4:k<4><<<148,512>>>(out,iters,1.f,cyc);break;case 5:k<5><<<148,512>>>(out,iters,1.f,cyc);break;} cudaDeviceSynchronize(); }
    printf("%-32s %6.2f cycles per iteration per scheduler\n", names[mode], (double)*cyc/iters/4.0); }
  printf("err=%s\n",cudaGetErrorString(cudaGetLastError()));
}
