// Exhaustive / randomised host check of the byte-parallel decimal formatting that the device P3
// writer is built from (raytracing-clj_b200/csrc/rtclj_p3_swar.h).  Built and run by tests/test_host.py.
#include "rtclj_p3_swar.h"
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <string>
using namespace rtclj;
int main(){
  // exhaustive digits
  for (unsigned a=0;a<256;++a) for (unsigned b=0;b<256;++b){
    P3Digits2 d=p3_digits2(a|(b<<16));
    unsigned ha=a/100,ta=(a/10)%10,oa=a%10,hb=b/100,tb=(b/10)%10,ob=b%10;
    if (d.ht!=((ha|(ta<<8))|((hb|(tb<<8))<<16)) || d.o!=(oa|(ob<<16))) {printf("digits bad %u %u\n",a,b);return 1;}
    uint32_t f0=p3_field(d,0,0x20), f1=p3_field(d,1,0x0a);
    if (f0!=((0x30+ha)|((0x30+ta)<<8)|((0x30+oa)<<16)|(0x20u<<24))) {printf("f0 bad\n");return 1;}
    if (f1!=((0x30+hb)|((0x30+tb)<<8)|((0x30+ob)<<16)|(0x0au<<24))) {printf("f1 bad\n");return 1;}
  }
  // extras / len
  srand(1);
  for (long it=0;it<4000000;++it){
    uint32_t w=((uint32_t)rand()<<16)^rand()^((uint32_t)rand()<<30);
    if (it<256*256) w = (it&0xff)|((it>>8)<<8)|0x63090a00u<<8;
    uint32_t e=p3_extra_digits4(w); unsigned len=0;
    for(int k=0;k<4;++k){unsigned v=(w>>(8*k))&0xff; unsigned x=(v>=10)+(v>=100); if(((e>>(8*k))&0xff)!=x){printf("extra bad %08x\n",w);return 1;} len+=2+x;}
    if (p3_len4(w)!=len){printf("len bad %08x\n",w);return 1;}
  }
  // thread emulation: 12 values, arbitrary start alignment
  for (long it=0;it<1000000;++it){
    uint32_t w[3]; unsigned char v[12];
    for(int k=0;k<12;++k){ int m=rand()%4; v[k]= m==0? rand()%10 : m==1? rand()%100 : rand()%256; }
    int n = (it%7==0)? 1+rand()%4 : 4;
    for(int k=3*n;k<12;++k) v[k]=0;
    memcpy(w,v,12);
    unsigned start=rand()%4;
    unsigned char buf[64]; memset(buf,0,64);
    uint32_t* b32=(uint32_t*)buf;
    P3Acc a; a.fill8=8*start; unsigned widx=0;
    for(int wi=0;wi<3;++wi){
      uint32_t e=p3_extra_digits4(w[wi]); uint32_t dz=0x10101010u-(e<<3);
      if(n<4) dz+=p3_padding_lanes(wi,n);
      P3Digits2 dl=p3_digits2(w[wi]&0x00ff00ffu), dh=p3_digits2((w[wi]>>8)&0x00ff00ffu);
      for(int j=0;j<4;++j){ int k=4*wi+j;
        uint32_t f=p3_field((j&1)?dh:dl, j>>1, (k%3==2)?0x0a:0x20);
        p3_acc_append(a,f,(dz>>(8*j))&0xff);
        if(p3_acc_full(a)) b32[widx++]|=p3_acc_pop(a);
      }
    }
    if(a.fill8) b32[widx]|=a.lo;
    std::string want; for(int k=0;k<3*n;++k){ char t[8]; sprintf(t,"%u%c",v[k],(k%3==2)?'\n':' '); want+=t; }
    unsigned len=p3_len4(w[0])+p3_len4(w[1])+p3_len4(w[2])-2*(12-3*n);
    if(len!=want.size()|| memcmp(buf+start,want.data(),want.size())!=0 ){printf("emul bad it=%ld n=%d start=%u len=%u want=%zu\n",it,n,start,len,want.size());return 1;}
    for(unsigned i=start+want.size();i<64;++i) if(buf[i]){printf("trailing garbage\n");return 1;}
    for(unsigned i=0;i<start;++i) if(buf[i]){printf("leading garbage\n");return 1;}
  }
  printf("ok\n");
}
