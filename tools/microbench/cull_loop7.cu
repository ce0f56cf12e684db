// Microbenchmark: how fast does the packed cull run with 1, 2 or 4 warps per scheduler culling, with the
// sphere table (0) in shared memory, (1) in constant memory, (2) half of the pairs from each?
#include <cstdio>
#include "../../raytracing-clj_b200/csrc/rtclj_kernels.cuh"
using namespace rtclj;
__constant__ uint4 ctab[2048];
template<int V> __global__ void __launch_bounds__(512,1) k(const float4* g, int nblocks, int reps, unsigned* out, long long* cyc){
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float4* sg=(float4*)smem_raw;
  for(int i=threadIdx.x;i<nblocks*16;i+=blockDim.x) sg[i]=g[i];
  __syncthreads();
  const unsigned smem_base=(unsigned)__cvta_generic_to_shared(smem_raw);
  float t=threadIdx.x*1e-3f;
  const f32x2 nbeta=splat2(-0.3f+t), kq=splat2(-1.5f-t), o2x=splat2(2.f*t), o2y=splat2(0.4f+t), o2z=splat2(-0.2f+t), dx2=splat2(0.6f+t), dy2=splat2(t), dz2=splat2(0.8f-t);
  unsigned total=0;
  long long t0=clock64();
  for(int r=0;r<reps;++r){
    unsigned addr=smem_base;
#pragma unroll 1
    for(int blk=0;blk<nblocks;++blk,addr+=256u){
      unsigned acc=0xffffffffu;
#pragma unroll
      for(int p=0;p<8;++p){
        f32x2 cx,cy,cz,rs;
        const bool from_smem = V==0 || (V==2 && (p&1));
        if (from_smem){ lds_pair(addr+32u*p,cx,cy); lds_pair(addr+32u*p+16u,cz,rs); }
        else { const uint4 a=ctab[(blk*8+p)*2], b=ctab[(blk*8+p)*2+1]; cx=((f32x2)a.y<<32)|a.x; cy=((f32x2)a.w<<32)|a.z; cz=((f32x2)b.y<<32)|b.x; rs=((f32x2)b.w<<32)|b.z; }
        const f32x2 bb=fma2(cz,dz2,fma2(cy,dy2,fma2(cx,dx2,nbeta)));
        const f32x2 ss=fma2(cz,o2z,fma2(cy,o2y,fma2(cx,o2x,add2(rs,kq))));
        const f32x2 dd=fma2(bb,bb,ss);
        acc=__funnelshift_l((unsigned)dd,acc,1); acc=__funnelshift_l((unsigned)(dd>>32),acc,1);
      }
      total+=acc;
    }
  }
  long long t1=clock64();
  out[blockIdx.x*blockDim.x+threadIdx.x]=total;
  if(threadIdx.x==0&&blockIdx.x==0) *cyc=t1-t0;
}
int main(){
  const int nblocks=31;
  float4* g; cudaMallocManaged(&g,nblocks*256);
  for(int i=0;i<nblocks*16;i++) g[i]=make_float4(5.f+i*0.01f,6.f,7.f+i*0.02f,-1e30f);
  cudaMemcpyToSymbol(ctab,g,nblocks*256);
  unsigned* out; long long* cyc; cudaMalloc(&out,148*512*4); cudaMallocManaged(&cyc,8);
  const int reps=2000;
  const char* names[3]={"shared (LDS.128)","constant (LDCU)","half and half"};
  for(int threads=128;threads<=512;threads*=2)
    for(int v=0;v<3;v++){
      for(int rep=0;rep<2;rep++){
        if(v==0) k<0><<<148,threads,nblocks*256>>>(g,nblocks,reps,out,cyc);
        else if(v==1) k<1><<<148,threads,nblocks*256>>>(g,nblocks,reps,out,cyc);
        else k<2><<<148,threads,nblocks*256>>>(g,nblocks,reps,out,cyc);
        cudaDeviceSynchronize(); }
      const int wps=threads/128;
      printf("%d warp(s)/scheduler  %-18s %6.2f cycles per pair per warp, %6.2f per scheduler (ideal 16)  (%s)\n", wps, names[v],
             (double)*cyc/reps/(nblocks*8), (double)*cyc/reps/(nblocks*8)/wps, cudaGetErrorString(cudaGetLastError())); }
}
