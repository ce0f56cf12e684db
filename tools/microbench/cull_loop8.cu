// Microbenchmark: R rays per lane sharing every sphere-pair load (R = 1, 2, 4), one warp per scheduler
// (and 2) culling, table in shared or constant memory.  Question: can ONE warp per scheduler keep the
// FFMA2 pipe busy when each load feeds R rays?
#include <cstdio>
#include "../../raytracing-clj_b200/csrc/rtclj_kernels.cuh"
using namespace rtclj;
__constant__ uint4 ctab[2048];
template<int V, int R> __global__ void __launch_bounds__(512,1) k(const float4* g, int nblocks, int reps, unsigned* out, long long* cyc){
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float4* sg=(float4*)smem_raw;
  for(int i=threadIdx.x;i<nblocks*16;i+=blockDim.x) sg[i]=g[i];
  __syncthreads();
  const unsigned smem_base=(unsigned)__cvta_generic_to_shared(smem_raw);
  f32x2 nbeta[R],kq[R],o2x[R],o2y[R],o2z[R],dx2[R],dy2[R],dz2[R];
#pragma unroll
  for(int r=0;r<R;++r){ float t=(threadIdx.x+97*r)*1e-3f; nbeta[r]=splat2(-0.3f+t); kq[r]=splat2(-1.5f-t); o2x[r]=splat2(2.f*t); o2y[r]=splat2(0.4f+t); o2z[r]=splat2(-0.2f+t); dx2[r]=splat2(0.6f+t); dy2[r]=splat2(t); dz2[r]=splat2(0.8f-t); }
  unsigned total=0;
  long long t0=clock64();
  for(int rep=0;rep<reps;++rep){
    unsigned addr=smem_base;
#pragma unroll 1
    for(int blk=0;blk<nblocks;++blk,addr+=256u){
      unsigned acc[R];
#pragma unroll
      for(int r=0;r<R;++r) acc[r]=0xffffffffu;
#pragma unroll
      for(int p=0;p<8;++p){
        f32x2 cx,cy,cz,rs;
        if (V==0){ lds_pair(addr+32u*p,cx,cy); lds_pair(addr+32u*p+16u,cz,rs); }
        else { const uint4 a=ctab[(blk*8+p)*2], b=ctab[(blk*8+p)*2+1]; cx=((f32x2)a.y<<32)|a.x; cy=((f32x2)a.w<<32)|a.z; cz=((f32x2)b.y<<32)|b.x; rs=((f32x2)b.w<<32)|b.z; }
#pragma unroll
        for(int r=0;r<R;++r){
          const f32x2 bb=fma2(cz,dz2[r],fma2(cy,dy2[r],fma2(cx,dx2[r],nbeta[r])));
          const f32x2 ss=fma2(cz,o2z[r],fma2(cy,o2y[r],fma2(cx,o2x[r],add2(rs,kq[r]))));
          const f32x2 dd=fma2(bb,bb,ss);
          acc[r]=__funnelshift_l((unsigned)dd,acc[r],1); acc[r]=__funnelshift_l((unsigned)(dd>>32),acc[r],1);
        }
      }
#pragma unroll
      for(int r=0;r<R;++r) total+=acc[r];
    }
  }
  long long t1=clock64();
  out[blockIdx.x*blockDim.x+threadIdx.x]=total;
  if(threadIdx.x==0&&blockIdx.x==0) *cyc=t1-t0;
}
template<int V,int R> void run(const float4* g,int nblocks,int reps,unsigned* out,long long* cyc,int threads){
  for(int rep=0;rep<2;rep++){ k<V,R><<<148,threads,nblocks*256>>>(g,nblocks,reps,out,cyc); cudaDeviceSynchronize(); }
  const int wps=threads/128;
  printf("%d warp/sched  %-9s %d ray(s)/lane: %6.2f cycles per (pair x ray) per scheduler (ideal 16)  (%s)\n", wps, V?"constant":"shared", R,
         (double)*cyc/reps/(nblocks*8)/wps/R, cudaGetErrorString(cudaGetLastError()));
}
int main(){
  const int nblocks=31;
  float4* g; cudaMallocManaged(&g,nblocks*256);
  for(int i=0;i<nblocks*16;i++) g[i]=make_float4(5.f+i*0.01f,6.f,7.f+i*0.02f,-1e30f);
  cudaMemcpyToSymbol(ctab,g,nblocks*256);
  unsigned* out; long long* cyc; cudaMalloc(&out,148*512*4); cudaMallocManaged(&cyc,8);
  const int reps=1000;
  for(int threads=128;threads<=256;threads*=2){
    run<0,1>(g,nblocks,reps,out,cyc,threads); run<0,2>(g,nblocks,reps,out,cyc,threads); run<0,4>(g,nblocks,reps,out,cyc,threads);
    run<1,1>(g,nblocks,reps,out,cyc,threads); run<1,2>(g,nblocks,reps,out,cyc,threads); run<1,4>(g,nblocks,reps,out,cyc,threads);
  }
}
