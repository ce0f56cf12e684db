// Microbenchmark: instruction arrangement of the packed cull (register-file / reuse-cache effects).
// All 16 warps cull.  Reports cycles per sphere pair per scheduler (8 FP2 per pair -> ideal 16).
#include <cstdio>
#include "../../raytracing-clj_b200/csrc/rtclj_kernels.cuh"
using namespace rtclj;
#define VFMA2(d,a,b,c) asm volatile("fma.rn.f32x2 %0, %1, %2, %3;":"=l"(d):"l"(a),"l"(b),"l"(c))
#define VADD2(d,a,b) asm volatile("add.rn.f32x2 %0, %1, %2;":"=l"(d):"l"(a),"l"(b))
__device__ __forceinline__ f32x2 hard_splat(float v){ f32x2 d; asm volatile("mov.b64 %0, {%1,%1};":"=l"(d):"f"(v)); return d; }
template<int V> __global__ void __launch_bounds__(512,1) k(const float4* g, int nblocks, int reps, unsigned* out, long long* cyc){
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float4* sg=(float4*)smem_raw;
  for(int i=threadIdx.x;i<nblocks*32;i+=blockDim.x) sg[i]=g[i];
  __syncthreads();
  const unsigned smem_base=(unsigned)__cvta_generic_to_shared(smem_raw);
  float t=threadIdx.x*1e-3f;
  f32x2 nbeta,kq,o2x,o2y,o2z,dx2,dy2,dz2;
  if (V==1){ nbeta=hard_splat(-0.3f+t); kq=hard_splat(-1.5f-t); o2x=hard_splat(2.f*t); o2y=hard_splat(0.4f+t); o2z=hard_splat(-0.2f+t); dx2=hard_splat(0.6f+t); dy2=hard_splat(t); dz2=hard_splat(0.8f-t); }
  else { nbeta=splat2(-0.3f+t); kq=splat2(-1.5f-t); o2x=splat2(2.f*t); o2y=splat2(0.4f+t); o2z=splat2(-0.2f+t); dx2=splat2(0.6f+t); dy2=splat2(t); dz2=splat2(0.8f-t); }
  unsigned total=0;
  long long t0=clock64();
  for(int r=0;r<reps;++r){
    unsigned addr=smem_base;
    for(int blk=0;blk<nblocks;++blk,addr+=512u){
      unsigned acc=0xffffffffu;
      if (V<=1){
#pragma unroll
        for(int p=0;p<16;++p){
          f32x2 cx,cy,cz,rs; lds_pair(addr+32u*p,cx,cy); lds_pair(addr+32u*p+16u,cz,rs);
          const f32x2 bb=fma2(cz,dz2,fma2(cy,dy2,fma2(cx,dx2,nbeta)));
          const f32x2 ss=fma2(cz,o2z,fma2(cy,o2y,fma2(cx,o2x,add2(rs,kq))));
          const f32x2 dd=fma2(bb,bb,ss);
          acc=__funnelshift_l((unsigned)dd,acc,1); acc=__funnelshift_l((unsigned)(dd>>32),acc,1);
        }
      } else {
        // stage-major, 8 pairs at a time, order pinned with volatile asm
#pragma unroll
        for(int h=0;h<2;++h){
          f32x2 cx[8],cy[8],cz[8],rs[8],bb[8],ss[8];
#pragma unroll
          for(int p=0;p<8;++p){ lds_pair(addr+256u*h+32u*p,cx[p],cy[p]); lds_pair(addr+256u*h+32u*p+16u,cz[p],rs[p]); }
#pragma unroll
          for(int p=0;p<8;++p) VADD2(ss[p],rs[p],kq);
#pragma unroll
          for(int p=0;p<8;++p) VFMA2(bb[p],cx[p],dx2,nbeta);
#pragma unroll
          for(int p=0;p<8;++p) VFMA2(ss[p],cx[p],o2x,ss[p]);
#pragma unroll
          for(int p=0;p<8;++p) VFMA2(bb[p],cy[p],dy2,bb[p]);
#pragma unroll
          for(int p=0;p<8;++p) VFMA2(ss[p],cy[p],o2y,ss[p]);
#pragma unroll
          for(int p=0;p<8;++p) VFMA2(bb[p],cz[p],dz2,bb[p]);
#pragma unroll
          for(int p=0;p<8;++p) VFMA2(ss[p],cz[p],o2z,ss[p]);
#pragma unroll
          for(int p=0;p<8;++p){ f32x2 dd; VFMA2(dd,bb[p],bb[p],ss[p]); acc=__funnelshift_l((unsigned)dd,acc,1); acc=__funnelshift_l((unsigned)(dd>>32),acc,1); }
        }
      }
      if(acc!=0xffffffffu) total+=__popc(~acc);
    }
  }
  long long t1=clock64();
  out[blockIdx.x*blockDim.x+threadIdx.x]=total;
  if(threadIdx.x==0&&blockIdx.x==0) *cyc=t1-t0;
}
template<int V> void run(const char* name, const float4* g, unsigned* out, long long* cyc){
  const int nblocks=16, reps=2000;
  cudaFuncSetAttribute(k<V>,cudaFuncAttributeMaxDynamicSharedMemorySize,100000);
  for(int rep=0;rep<2;rep++){ k<V><<<148,512,nblocks*512>>>(g,nblocks,reps,out,cyc); cudaDeviceSynchronize(); }
  printf("%-44s %5.2f   (%s)\n", name, (double)*cyc/reps/(nblocks*16)/4, cudaGetErrorString(cudaGetLastError()));
}
int main(){
  float4* g; cudaMallocManaged(&g,16*512);
  for(int i=0;i<16*32;i++) g[i]=make_float4(5.f+i*0.01f,6.f,7.f+i*0.02f,-1e30f);
  unsigned* out; long long* cyc; cudaMalloc(&out,148*512*4); cudaMallocManaged(&cyc,8);
  printf("cycles per sphere pair per scheduler, 4 warps culling (ideal 16)\n");
  run<0>("V0 compiler order, scalar-broadcast consts", g,out,cyc);
  run<1>("V1 compiler order, 64-bit splat consts", g,out,cyc);
  run<2>("V2 stage-major order (8 pairs), pinned", g,out,cyc);
}
