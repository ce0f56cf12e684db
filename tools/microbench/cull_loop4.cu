// Microbenchmark: what limits a LONE warp in the packed cull loop?  Variants drop the sign
// shifts and/or the shared-memory loads.  1 or 4 of a scheduler's 4 warps cull; the rest run
// a dependent fp64/int chain.
#include <cstdio>
#include "../../raytracing-clj_b200/csrc/rtclj_kernels.cuh"
using namespace rtclj;
template<int SIGN, int LOADS> __global__ void __launch_bounds__(512,1) k(const float4* g, int nblocks, int reps, unsigned* out, long long* cyc, int ncull){
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float4* sg=(float4*)smem_raw;
  for(int i=threadIdx.x;i<nblocks*32;i+=blockDim.x) sg[i]=g[i];
  __syncthreads();
  const unsigned smem_base=(unsigned)__cvta_generic_to_shared(smem_raw);
  float t=threadIdx.x*1e-3f;
  const f32x2 nbeta=splat2(-0.3f+t), kq=splat2(-1.5f-t);
  const f32x2 o2x=splat2(2.f*t), o2y=splat2(0.4f+t), o2z=splat2(-0.2f+t);
  const f32x2 dx2=splat2(0.6f), dy2=splat2(0.0f+t), dz2=splat2(0.8f);
  unsigned total=0;
  const int warp=threadIdx.x>>5;
  if ((warp>>2) >= ncull) {
    double x=1.0+t, y=0.5; unsigned n=threadIdx.x;
    for(int r=0;r<reps*nblocks*12;++r){ x=x*y+0.25; y=y/(x+1.0); n=n*1664525u+1013904223u; if(n&1) x+=1e-3; }
    out[blockIdx.x*blockDim.x+threadIdx.x]=(unsigned)x+n; return;
  }
  long long t0=clock64();
  for(int r=0;r<reps;++r){
    unsigned addr=smem_base;
    for(int blk=0;blk<nblocks;++blk,addr+=512u){
      unsigned acc=0xffffffffu;
      f32x2 cx,cy,cz,rs;
#pragma unroll
      for(int p=0;p<16;++p){
        if (LOADS==2 || (LOADS==1 && (p&1)==0) || (LOADS==0 && p==0)) { lds_pair(addr+32u*p,cx,cy); lds_pair(addr+32u*p+16u,cz,rs); }
        const f32x2 bb=fma2(cz,dz2,fma2(cy,dy2,fma2(cx,dx2,nbeta)));
        const f32x2 ss=fma2(cz,o2z,fma2(cy,o2y,fma2(cx,o2x,add2(rs,kq))));
        const f32x2 dd=fma2(bb,bb,ss);
        if (SIGN==2){ acc=__funnelshift_l((unsigned)dd,acc,1); acc=__funnelshift_l((unsigned)(dd>>32),acc,1); }
        else if (SIGN==1){ acc &= (unsigned)dd & (unsigned)(dd>>32); }
        else { if (p==15) acc&=(unsigned)dd; else { cx = cx ^ (dd & 1ull); } }
      }
      if(acc!=0xffffffffu && (int)acc>=0) total++;
    }
  }
  long long t1=clock64();
  out[blockIdx.x*blockDim.x+threadIdx.x]=total;
  if(threadIdx.x==0&&blockIdx.x==0) *cyc=t1-t0;
}
template<int SIGN,int LOADS> void run(const char* name, const float4* g, unsigned* out, long long* cyc){
  const int nblocks=16, reps=2000;
  cudaFuncSetAttribute(k<SIGN,LOADS>,cudaFuncAttributeMaxDynamicSharedMemorySize,100000);
  printf("%-44s", name);
  for(int ncull: {4,1}){ for(int rep=0;rep<2;rep++){ k<SIGN,LOADS><<<148,512,nblocks*512>>>(g,nblocks,reps,out,cyc,ncull); cudaDeviceSynchronize(); }
    printf("  %d/4: %5.2f", ncull, (double)*cyc/reps/(nblocks*16)/ncull); }
  printf("   (%s)\n", cudaGetErrorString(cudaGetLastError()));
}
int main(){
  float4* g; cudaMallocManaged(&g,16*512);
  for(int i=0;i<16*32;i++) g[i]=make_float4(5.f+i*0.01f,6.f,7.f+i*0.02f,-1e30f);
  unsigned* out; long long* cyc; cudaMalloc(&out,148*512*4); cudaMallocManaged(&cyc,8);
  printf("cycles per sphere pair per scheduler (ideal 16)\n");
  run<2,2>("2 SHF + 2 LDS per pair (the kernel's loop)", g,out,cyc);
  run<1,2>("1 LOP3 + 2 LDS per pair", g,out,cyc);
  run<0,2>("no sign op + 2 LDS per pair", g,out,cyc);
  run<2,1>("2 SHF + 1 LDS per pair", g,out,cyc);
  run<2,0>("2 SHF + loads once per block", g,out,cyc);
  run<0,0>("FP2 only", g,out,cyc);
}
