// Microbenchmark: cost of FFMA2 by operand pattern (register-file bandwidth / reuse cache).
// 16 warps per SM (4 per scheduler); prints cycles per FFMA2 per scheduler.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
#define FMA2(d,a,b,c) asm volatile("fma.rn.f32x2 %0, %1, %2, %3;":"=l"(d):"l"(a),"l"(b),"l"(c))
__device__ __forceinline__ u64 splat(float v){u64 d; asm("mov.b64 %0, {%1,%1};":"=l"(d):"f"(v)); return d;}
template<int MODE> __global__ void __launch_bounds__(512) k(float* out, int iters, float seed, long long* cyc){
  u64 p[8], acc[8];
  for(int i=0;i<8;i++){ p[i]=splat(seed+i+threadIdx.x*1e-3f); acc[i]=splat(seed*0.5f+i); }
  float s1=seed+threadIdx.x*1e-4f, s2=seed*2+threadIdx.x*1e-4f, s3=seed*3+threadIdx.x*1e-4f, s4 = 1e-9f+threadIdx.x*1e-12f;
  u64 S1=splat(s1), S2=splat(s2), S3=splat(s3), S4=splat(s4);
  long long t0=clock64();
  for(int it=0;it<iters;++it){
    if(MODE==0){ _Pragma("unroll") for(int i=0;i<8;i++) FMA2(acc[i],p[i],S1,S4); }          // pair, scalar, scalar -> new dst
    if(MODE==1){ _Pragma("unroll") for(int i=0;i<8;i++) FMA2(acc[i],p[i],S1,acc[i]); }       // pair, scalar(same), pair(acc)
    if(MODE==2){ _Pragma("unroll") for(int i=0;i<8;i++) FMA2(acc[i],p[i],(i%3==0?S1:(i%3==1?S2:S3)),acc[i]); } // rotating scalars
    if(MODE==3){ _Pragma("unroll") for(int i=0;i<8;i++) FMA2(acc[i],p[i],p[(i+3)&7],acc[i]); } // three pairs
    if(MODE==4){ _Pragma("unroll") for(int i=0;i<8;i++) FMA2(acc[i],acc[i],acc[i],p[i]); }     // b*b + pair
    if(MODE==5){ // the cull's pattern: bb=fma(cz,dz,fma(cy,dy,fma(cx,dx,nb))); ss similarly; dd=fma(bb,bb,ss)
      _Pragma("unroll") for(int i=0;i<8;i+=4){
        u64 b,s,d; FMA2(b,p[i],S1,S4); FMA2(b,p[i+1],S2,b); FMA2(b,p[i+2],S3,b);
        u64 w; asm volatile("add.rn.f32x2 %0, %1, %2;":"=l"(w):"l"(p[i+3]),"l"(S4));
        FMA2(s,p[i],S3,w); FMA2(s,p[i+1],S1,s); FMA2(s,p[i+2],S2,s); FMA2(d,b,b,s); acc[i]^=d; }
    }
  }
  long long t1=clock64();
  float s=0; for(int i=0;i<8;i++){ s+=__uint_as_float((unsigned)acc[i])+__uint_as_float((unsigned)(acc[i]>>32)); }
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
  if(threadIdx.x==0&&blockIdx.x==0) *cyc=t1-t0;
}
int main(){
  float* out; long long* cyc; cudaMalloc(&out,148*512*4); cudaMallocManaged(&cyc,8);
  const int iters=20000; const char* names[]={"pair,scalar,scalar","pair,scalar(same),acc","pair,scalar(rot3),acc","pair,pair,acc","acc*acc+pair","cull pattern (16 FP2/iter)"};
  for(int mode=0;mode<6;mode++){
    for(int rep=0;rep<2;rep++){
      switch(mode){case 0:k<0><<<148,512>>>(out,iters,1.f,cyc);break;case 1:k<1><<<148,512>>>(out,iters,1.f,cyc);break;case 2:k<2><<<148,512>>>(out,iters,1.f,cyc);break;case 3:k<3><<<148,512>>>(out,iters,1.f,cyc);break;case 4:k<4><<<148,512>>>(out,iters,1.f,cyc);break;case 5:k<5><<<148,512>>>(out,iters,1.f,cyc);break;}
      cudaDeviceSynchronize();
    }
    const int n = mode==5?16:8;
    printf("%-28s %6.2f cycles per FFMA2 per scheduler\n", names[mode], (double)*cyc/iters/4.0/n);
  }
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
}
