// Microbenchmark: does FFMA2 (packed fp32 FMA) occupy the issue port for 2 cycles?
// Runs per-thread loops of (a) 8 FFMA2, (b) 8 FFMA2 + 8 independent integer ops,
// (c) 8 FFMA2 + 16 integer ops, (d) 16 scalar FFMA, (e) 16 FFMA + 8 int, (f) 8 FFMA2 + 8 LDS.
// Prints cycles per loop iteration per SM sub-partition (4 warps resident per scheduler).
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c){u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;":"=l"(d):"l"(a),"l"(b),"l"(c)); return d;}
template<int MODE> __global__ void __launch_bounds__(512) k(float* out, int iters, unsigned seed, long long* cyc){
  __shared__ float sm[1024];
  sm[threadIdx.x] = threadIdx.x; sm[threadIdx.x+512]=1.f; __syncthreads();
  u64 a[8]; float f[16]; unsigned n[16];
  for(int i=0;i<8;i++){ a[i] = ((u64)__float_as_uint(1.0f+i)<<32)|__float_as_uint(2.0f+i);} 
  for(int i=0;i<16;i++){ f[i]=1.0f+i; n[i]=seed+i+threadIdx.x; }
  const u64 m = ((u64)__float_as_uint(0.9999f)<<32)|__float_as_uint(0.9999f), c=((u64)__float_as_uint(1e-9f)<<32)|__float_as_uint(1e-9f);
  float ls=0.f;
  long long t0 = clock64();
  for(int it=0; it<iters; ++it){
    if (MODE<=2 || MODE==5){
      #pragma unroll
      for(int i=0;i<8;i++) a[i]=fma2(a[i],m,c);
    } else {
      #pragma unroll
      for(int i=0;i<16;i++) asm volatile("fma.rn.f32 %0, %0, %1, %2;":"+f"(f[i]):"f"(0.9999f),"f"(1e-9f));
    }
    if (MODE==1||MODE==2||MODE==4){
      #pragma unroll
      for(int i=0;i<8;i++) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;":"+r"(n[i]):"r"(seed),"r"(n[(i+1)&7]));
    }
    if (MODE==2){
      #pragma unroll
      for(int i=8;i<16;i++) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;":"+r"(n[i]):"r"(seed),"r"(n[8+((i+1)&7)]));
    }
    if (MODE==5){
      #pragma unroll
      for(int i=0;i<8;i++){ float v; asm volatile("ld.shared.f32 %0, [%1];":"=f"(v):"r"((unsigned)__cvta_generic_to_shared(sm + ((it*8+i)&1023)))); ls+=v; }
    }
  }
  long long t1 = clock64();
  float s=ls; for(int i=0;i<8;i++){ s+=__uint_as_float((unsigned)a[i])+__uint_as_float((unsigned)(a[i]>>32)); }
  for(int i=0;i<16;i++) s+=f[i]+(float)n[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
  if (threadIdx.x==0 && blockIdx.x==0) *cyc = t1-t0;
}
int main(){
  float* out; long long* cyc; cudaMalloc(&out, 148*512*4); cudaMallocManaged(&cyc, 8);
  const int iters=20000; const char* names[]={"8xFFMA2","8xFFMA2+8xLOP3","8xFFMA2+16xLOP3","16xFFMA","16xFFMA+8xLOP3","8xFFMA2+8x(LDS+FADD)"};
  for(int mode=0;mode<6;mode++){
    for(int rep=0;rep<2;rep++){
      switch(mode){case 0:k<0><<<148,512>>>(out,iters,1,cyc);break;case 1:k<1><<<148,512>>>(out,iters,1,cyc);break;case 2:k<2><<<148,512>>>(out,iters,1,cyc);break;case 3:k<3><<<148,512>>>(out,iters,1,cyc);break;case 4:k<4><<<148,512>>>(out,iters,1,cyc);break;case 5:k<5><<<148,512>>>(out,iters,1,cyc);break;}
      cudaDeviceSynchronize();
    }
    // 512 threads = 16 warps = 4 warps per scheduler; cycles per iteration per scheduler-warp
    printf("%-24s %8.2f cycles/iter/block -> %6.2f cycles per warp-iteration per scheduler\n", names[mode], (double)*cyc/iters, (double)*cyc/iters/4.0);
  }
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
