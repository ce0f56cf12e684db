// Microbenchmark: sphere table in CONSTANT memory (uniform loads, no per-lane LDS) vs shared memory.
#include <cstdio>
#include "../../raytracing-clj_b200/csrc/rtclj_kernels.cuh"
using namespace rtclj;
__constant__ uint4 ctab[2048];
template<int V> __global__ void __launch_bounds__(512,1) k(const float4* g, int nblocks, int reps, unsigned* out, long long* cyc){
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float4* sg=(float4*)smem_raw;
  for(int i=threadIdx.x;i<nblocks*32;i+=blockDim.x) sg[i]=g[i];
  __syncthreads();
  const unsigned smem_base=(unsigned)__cvta_generic_to_shared(smem_raw);
  float t=threadIdx.x*1e-3f;
  const f32x2 nbeta=splat2(-0.3f+t), kq=splat2(-1.5f-t), o2x=splat2(2.f*t), o2y=splat2(0.4f+t), o2z=splat2(-0.2f+t), dx2=splat2(0.6f+t), dy2=splat2(t), dz2=splat2(0.8f-t);
  unsigned total=0;
  long long t0=clock64();
  for(int r=0;r<reps;++r){
    unsigned addr=smem_base;
    for(int blk=0;blk<nblocks;++blk,addr+=512u){
      unsigned acc=0xffffffffu;
#pragma unroll
      for(int p=0;p<16;++p){
        f32x2 cx,cy,cz,rs;
        if (V==0){ lds_pair(addr+32u*p,cx,cy); lds_pair(addr+32u*p+16u,cz,rs); }
        else { const uint4 a=ctab[(blk*16+p)*2], b=ctab[(blk*16+p)*2+1]; cx=((f32x2)a.y<<32)|a.x; cy=((f32x2)a.w<<32)|a.z; cz=((f32x2)b.y<<32)|b.x; rs=((f32x2)b.w<<32)|b.z; }
        const f32x2 bb=fma2(cz,dz2,fma2(cy,dy2,fma2(cx,dx2,nbeta)));
        const f32x2 ss=fma2(cz,o2z,fma2(cy,o2y,fma2(cx,o2x,add2(rs,kq))));
        const f32x2 dd=fma2(bb,bb,ss);
        acc=__funnelshift_l((unsigned)dd,acc,1); acc=__funnelshift_l((unsigned)(dd>>32),acc,1);
      }
      if(acc!=0xffffffffu) total+=__popc(~acc);
    }
  }
  long long t1=clock64();
  out[blockIdx.x*blockDim.x+threadIdx.x]=total;
  if(threadIdx.x==0&&blockIdx.x==0) *cyc=t1-t0;
}
int main(){
  const int nblocks=16;
  float4* g; cudaMallocManaged(&g,nblocks*512);
  for(int i=0;i<nblocks*32;i++) g[i]=make_float4(5.f+i*0.01f,6.f,7.f+i*0.02f,-1e30f);
  cudaMemcpyToSymbol(ctab,g,nblocks*512);
  unsigned* out; long long* cyc; cudaMalloc(&out,148*512*4); cudaMallocManaged(&cyc,8);
  const int reps=2000;
  for(int v=0;v<2;v++){ for(int rep=0;rep<2;rep++){ if(v==0) k<0><<<148,512,nblocks*512>>>(g,nblocks,reps,out,cyc); else k<1><<<148,512,nblocks*512>>>(g,nblocks,reps,out,cyc); cudaDeviceSynchronize(); }
    printf("%-28s %5.2f cycles per sphere pair per scheduler (ideal 16)  (%s)\n", v==0?"shared memory (LDS.128)":"constant memory (uniform)", (double)*cyc/reps/(nblocks*16)/4, cudaGetErrorString(cudaGetLastError())); }
}
