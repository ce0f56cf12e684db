// Microbenchmark: SCALAR-FFMA cull loop (one sphere per LDS.128, 8 scalar FMA-pipe ops per
// sphere) vs the packed FFMA2 loop, with 1..4 of a scheduler's 4 warps culling.
#include <cstdio>
#include "../../raytracing-clj_b200/csrc/rtclj_kernels.cuh"
using namespace rtclj;
__device__ __forceinline__ float4 lds4(unsigned a){ float4 v; asm("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];":"=f"(v.x),"=f"(v.y),"=f"(v.z),"=f"(v.w):"r"(a)); return v; }
template<int BS, int SIGN> __global__ void __launch_bounds__(512,1) k(const float4* g, int nblocks, int reps, unsigned* out, long long* cyc, int ncull){
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float4* sg=(float4*)smem_raw;
  for(int i=threadIdx.x;i<nblocks*BS;i+=blockDim.x) sg[i]=g[i];
  __syncthreads();
  const unsigned smem_base=(unsigned)__cvta_generic_to_shared(smem_raw);
  float t=threadIdx.x*1e-3f;
  const float nbeta=-0.3f+t, kq=-1.5f-t, o2x=2.f*t, o2y=0.4f+t, o2z=-0.2f+t, dx=0.6f, dy=t, dz=0.8f;
  unsigned total=0;
  const int warp=threadIdx.x>>5;
  if ((warp>>2) >= ncull) {
    double x=1.0+t, y=0.5; unsigned n=threadIdx.x;
    for(int r=0;r<reps*nblocks*BS*3/8;++r){ x=x*y+0.25; y=y/(x+1.0); n=n*1664525u+1013904223u; if(n&1) x+=1e-3; }
    out[blockIdx.x*blockDim.x+threadIdx.x]=(unsigned)x+n; return;
  }
  long long t0=clock64();
  for(int r=0;r<reps;++r){
    unsigned addr=smem_base;
    for(int blk=0;blk<nblocks;++blk,addr+=16u*BS){
      unsigned acc=0xffffffffu; float mx=-1e30f;
#pragma unroll
      for(int p=0;p<BS;++p){
        const float4 c=lds4(addr+16u*p);
        const float bb=fmaf(c.z,dz,fmaf(c.y,dy,fmaf(c.x,dx,nbeta)));
        const float ss=fmaf(c.z,o2z,fmaf(c.y,o2y,fmaf(c.x,o2x,c.w+kq)));
        const float dd=fmaf(bb,bb,ss);
        if (SIGN==0) acc=__funnelshift_l(__float_as_uint(dd),acc,1);
        else if (SIGN==1) acc&=__float_as_uint(dd);
        else mx=fmaxf(mx,dd);
      }
      if (SIGN==0){ if(acc!=0xffffffffu) total+=__popc(~acc); }
      else if (SIGN==1){ if((int)acc>=0) total++; }
      else { if(mx>=0.f) total++; }
    }
  }
  long long t1=clock64();
  out[blockIdx.x*blockDim.x+threadIdx.x]=total;
  if(threadIdx.x==0&&blockIdx.x==0) *cyc=t1-t0;
}
template<int BS,int SIGN> void run(const char* name, const float4* g, int nsph, unsigned* out, long long* cyc){
  const int nblocks=nsph/BS, reps=2000;
  cudaFuncSetAttribute(k<BS,SIGN>,cudaFuncAttributeMaxDynamicSharedMemorySize,100000);
  printf("%-34s", name);
  for(int ncull=4;ncull>=1;ncull--){ for(int rep=0;rep<2;rep++){ k<BS,SIGN><<<148,512,nsph*16>>>(g,nblocks,reps,out,cyc,ncull); cudaDeviceSynchronize(); }
    printf("  %d/4: %5.2f", ncull, 2.0*(double)*cyc/reps/(nblocks*BS)/ncull); }
  printf("   (%s)\n", cudaGetErrorString(cudaGetLastError()));
}
int main(){
  const int nsph=512;
  float4* g; cudaMallocManaged(&g,nsph*16);
  for(int i=0;i<nsph;i++) g[i]=make_float4(5.f+i*0.01f,6.f,7.f+i*0.02f,-1e30f);
  unsigned* out; long long* cyc; cudaMalloc(&out,148*512*4); cudaMallocManaged(&cyc,8);
  printf("scalar cull: cycles per sphere PAIR per scheduler (ideal 16)\n");
  run<16,0>("scalar, block 16, SHF sign", g,nsph,out,cyc);
  run<16,1>("scalar, block 16, LOP3 and", g,nsph,out,cyc);
  run<16,2>("scalar, block 16, FMNMX max", g,nsph,out,cyc);
  run<32,0>("scalar, block 32, SHF sign", g,nsph,out,cyc);
}
