// Microbenchmark: software-pipelined cull loop (prefetch K pairs ahead, survivor branch
// lagging one block) vs the plain block loop, with 1..4 of a scheduler's 4 warps culling.
#include <cstdio>
#include "../../raytracing-clj_b200/csrc/rtclj_kernels.cuh"
using namespace rtclj;
template<int BP, int K, bool LAG> __global__ void __launch_bounds__(512,1) k(const float4* g, int nblocks, int reps, unsigned* out, long long* cyc, int ncull){
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float4* sg=(float4*)smem_raw;
  for(int i=threadIdx.x;i<(nblocks*BP+K)*2;i+=blockDim.x) sg[i]=g[i];
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
    for(int r=0;r<reps*nblocks*BP*3/4;++r){ x=x*y+0.25; y=y/(x+1.0); n=n*1664525u+1013904223u; if(n&1) x+=1e-3; }
    out[blockIdx.x*blockDim.x+threadIdx.x]=(unsigned)x+n; return;
  }
  long long t0=clock64();
  for(int r=0;r<reps;++r){
    unsigned addr=smem_base;
    f32x2 bx[K>0?K:1], by[K>0?K:1], bz[K>0?K:1], bw[K>0?K:1];
#pragma unroll
    for(int q=0;q<K;++q){ lds_pair(addr+32u*q,bx[q],by[q]); lds_pair(addr+32u*q+16u,bz[q],bw[q]); }
    unsigned acc_prev=0xffffffffu;
    for(int blk=0;blk<nblocks;++blk,addr+=32u*BP){
      unsigned acc=0xffffffffu;
#pragma unroll
      for(int p=0;p<BP;++p){
        f32x2 cx,cy,cz,rs;
        if (K>0){ cx=bx[p%K]; cy=by[p%K]; cz=bz[p%K]; rs=bw[p%K];
          lds_pair(addr+32u*(p+K),bx[p%K],by[p%K]); lds_pair(addr+32u*(p+K)+16u,bz[p%K],bw[p%K]); }
        else { lds_pair(addr+32u*p,cx,cy); lds_pair(addr+32u*p+16u,cz,rs); }
        const f32x2 bb=fma2(cz,dz2,fma2(cy,dy2,fma2(cx,dx2,nbeta)));
        const f32x2 ss=fma2(cz,o2z,fma2(cy,o2y,fma2(cx,o2x,add2(rs,kq))));
        const f32x2 dd=fma2(bb,bb,ss);
        acc=__funnelshift_l((unsigned)dd,acc,1); acc=__funnelshift_l((unsigned)(dd>>32),acc,1);
        if (LAG && p==BP/2) { if(acc_prev!=0xffffffffu) total+=__popc(~acc_prev); }
      }
      if (LAG) acc_prev=acc; else { if(acc!=0xffffffffu) total+=__popc(~acc); }
    }
    if (LAG && acc_prev!=0xffffffffu) total+=__popc(~acc_prev);
  }
  long long t1=clock64();
  out[blockIdx.x*blockDim.x+threadIdx.x]=total;
  if(threadIdx.x==0&&blockIdx.x==0) *cyc=t1-t0;
}
template<int BP,int K,bool LAG> void run(const char* name, const float4* g, int npairs_total, unsigned* out, long long* cyc){
  const int nblocks=npairs_total/BP, reps=2000;
  cudaFuncSetAttribute(k<BP,K,LAG>,cudaFuncAttributeMaxDynamicSharedMemorySize,100000);
  printf("%-34s", name);
  for(int ncull=4;ncull>=1;ncull--){ for(int rep=0;rep<2;rep++){ k<BP,K,LAG><<<148,512,(npairs_total+16)*32>>>(g,nblocks,reps,out,cyc,ncull); cudaDeviceSynchronize(); }
    printf("  %d/4: %5.2f", ncull, (double)*cyc/reps/(nblocks*BP)/ncull); }
  printf("   (%s)\n", cudaGetErrorString(cudaGetLastError()));
}
int main(){
  const int npairs=256;
  float4* g; cudaMallocManaged(&g,(npairs+16)*32);
  for(int i=0;i<(npairs+16)*2;i++) g[i]=make_float4(5.f+i*0.01f,6.f,7.f+i*0.02f,3.f);
  for(int i=1;i<(npairs+16)*2;i+=2){ g[i].z=-1e30f; g[i].w=-1e30f; }
  unsigned* out; long long* cyc; cudaMalloc(&out,148*512*4); cudaMallocManaged(&cyc,8);
  printf("cycles per sphere pair per scheduler (ideal 16), by number of culling warps per scheduler\n");
  run<8,0,false>("block 8, no prefetch", g,npairs,out,cyc);
  run<8,0,true >("block 8, lagged branch", g,npairs,out,cyc);
  run<8,2,true >("block 8, prefetch 2, lagged", g,npairs,out,cyc);
  run<8,4,true >("block 8, prefetch 4, lagged", g,npairs,out,cyc);
  run<16,0,false>("block 16, no prefetch", g,npairs,out,cyc);
  run<16,4,true >("block 16, prefetch 4, lagged", g,npairs,out,cyc);
  run<16,4,false>("block 16, prefetch 4", g,npairs,out,cyc);
  run<16,8,true >("block 16, prefetch 8, lagged", g,npairs,out,cyc);
}
