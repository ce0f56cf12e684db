// Microbenchmark: the packed cull block loop in isolation (all 16 warps of every SM run it).
// Reports heavy-pipe cycles per sphere pair per scheduler; the ideal is 16 (8 FFMA2 x 2 cycles).
#include <cstdio>
#include "../../raytracing-clj_b200/csrc/rtclj_kernels.cuh"
using namespace rtclj;
// ncull = how many of each scheduler's 4 warps run the cull; the others run a dependent
// fp64 / integer chain (a stand-in for the shading phases) until the cull warps finish.
template<int VARIANT> __global__ void __launch_bounds__(512,1) k(const float4* g, int nblocks, int reps, unsigned* out, long long* cyc, int ncull){
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float4* sg=(float4*)smem_raw;
  for(int i=threadIdx.x;i<nblocks*kBlockPairs*2;i+=blockDim.x) sg[i]=g[i];
  __syncthreads();
  const unsigned smem_base=(unsigned)__cvta_generic_to_shared(smem_raw);
  float t=threadIdx.x*1e-3f;
  const f32x2 nbeta=splat2(-0.3f+t), kq=splat2(-1.5f-t);
  const f32x2 o2x=splat2(2.f*t), o2y=splat2(0.4f+t), o2z=splat2(-0.2f+t);
  const f32x2 dx2=splat2(0.6f), dy2=splat2(0.0f+t), dz2=splat2(0.8f);
  unsigned total=0;
  const int warp=threadIdx.x>>5;
  if ((warp>>2) >= ncull) {  // warps 4q..4q+3 sit on schedulers 0..3: warp>>2 = slot on its scheduler
    double x=1.0+t, y=0.5; unsigned n=threadIdx.x; 
    for(int r=0;r<reps*nblocks*6;++r){ x=x*y+0.25; y=y/(x+1.0); n=n*1664525u+1013904223u; if(n&1) x+=1e-3; }
    out[blockIdx.x*blockDim.x+threadIdx.x]=(unsigned)x+n; return;
  }
  long long t0=clock64();
  for(int r=0;r<reps;++r){
    unsigned addr=smem_base;
    for(int blk=0;blk<nblocks;++blk,addr+=32u*kBlockPairs){
      unsigned acc=0xffffffffu;
#pragma unroll
      for(int p=0;p<kBlockPairs;++p){
        f32x2 cx,cy,cz,rs; lds_pair(addr+32u*p,cx,cy); lds_pair(addr+32u*p+16u,cz,rs);
        const f32x2 bb=fma2(cz,dz2,fma2(cy,dy2,fma2(cx,dx2,nbeta)));
        const f32x2 ss=fma2(cz,o2z,fma2(cy,o2y,fma2(cx,o2x,add2(rs,kq))));
        const f32x2 dd=fma2(bb,bb,ss);
        if (VARIANT==0){ acc=__funnelshift_l((unsigned)dd,acc,1); acc=__funnelshift_l((unsigned)(dd>>32),acc,1); }
        else { acc &= (unsigned)dd & (unsigned)(dd>>32); }
      }
      if (VARIANT==0) { if(acc!=0xffffffffu) total+=__popc(~acc); }
      else { if ((int)acc>=0) total++; }
    }
  }
  long long t1=clock64();
  out[blockIdx.x*blockDim.x+threadIdx.x]=total;
  if(threadIdx.x==0&&blockIdx.x==0) *cyc=t1-t0;
}
int main(){
  const int nblocks=31; const int npairs=nblocks*kBlockPairs; const int reps=2000;
  float4* g; cudaMallocManaged(&g,npairs*32);
  for(int i=0;i<npairs*2;i++) g[i]=make_float4(5.f+i*0.01f,6.f,7.f+i*0.02f, (i&1)?-1e30f:3.f);
  for(int i=1;i<npairs*2;i+=2){ g[i].z=-1e30f; g[i].w=-1e30f; }
  unsigned* out; long long* cyc; cudaMalloc(&out,148*512*4); cudaMallocManaged(&cyc,8);
  cudaFuncSetAttribute(k<0>,cudaFuncAttributeMaxDynamicSharedMemorySize,100000);
  cudaFuncSetAttribute(k<1>,cudaFuncAttributeMaxDynamicSharedMemorySize,100000);
  for(int ncull=4;ncull>=1;ncull--){ for(int rep=0;rep<2;rep++){ k<0><<<148,512,npairs*32>>>(g,nblocks,reps,out,cyc,ncull); cudaDeviceSynchronize(); }
    printf("%d of 4 warps per scheduler culling: %.2f cycles per sphere pair per cull warp, %.2f per scheduler (ideal 16)\n", ncull, (double)*cyc/reps/npairs, (double)*cyc/reps/npairs/ncull); }
  printf("err=%s\n",cudaGetErrorString(cudaGetLastError()));
}
