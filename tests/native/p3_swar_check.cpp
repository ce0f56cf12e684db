// Exhaustive / randomised host check of the byte-parallel decimal formatting that the device P3
// writer is built from (raytracing-clj_b200/csrc/rtclj_p3_swar.h).  Built and run by tests/test_host.py.
#include "rtclj_p3_swar.h"
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <string>
using namespace rtclj;
int main(){
  // exhaustive digits and fields over all value pairs
  for (unsigned a=0;a<256;++a) for (unsigned b=0;b<256;++b){
    P3Digits2 d=p3_digits2(a|(b<<16));
    unsigned ha=a/100,ta=(a/10)%10,oa=a%10,hb=b/100,tb=(b/10)%10,ob=b%10;
    if (d.ht!=((ha|(ta<<8))|((hb|(tb<<8))<<16)) || d.o!=(oa|(ob<<16))) {printf("digits bad %u %u\n",a,b);return 1;}
    uint32_t f0=p3_field(d,0,0x20), f1=p3_field(d,1,0x0a);
    if (f0!=((0x30+ha)|((0x30+ta)<<8)|((0x30+oa)<<16)|(0x20u<<24))) {printf("f0 bad\n");return 1;}
    if (f1!=((0x30+hb)|((0x30+tb)<<8)|((0x30+ob)<<16)|(0x0au<<24))) {printf("f1 bad\n");return 1;}
  }
  // digit counts / lengths
  srand(1);
  for (long it=0;it<4000000;++it){
    uint32_t w=((uint32_t)rand()<<16)^rand()^((uint32_t)rand()<<30);
    if (it<256*256) w = (it&0xff)|((it>>8)<<8)|0x63090a00u<<8;
    uint32_t e=p3_extra_digits4(w); unsigned len=0;
    for(int k=0;k<4;++k){unsigned v=(w>>(8*k))&0xff; unsigned x=(v>=10)+(v>=100); if(((e>>(8*k))&0xff)!=x){printf("extra bad %08x\n",w);return 1;} len+=2+x;}
    if (p3_len4(w)!=len){printf("len bad %08x\n",w);return 1;}
  }
  // one pixel: every digit-count combination and random values, against sprintf
  for (long it=0;it<2000000;++it){
    unsigned v[3];
    for(int k=0;k<3;++k){ int m=(it>>(2*k))&3; v[k]= m==0? rand()%10 : m==1? 10+rand()%90 : m==2? 100+rand()%156 : rand()%256; }
    uint32_t w=v[0]|(v[1]<<8)|(v[2]<<16);
    uint32_t drop=0x10101010u-(p3_extra_digits4(w)<<3);
    P3Digits2 even=p3_digits2(w&0x00ff00ffu), odd=p3_digits2((w>>8)&0x00ff00ffu);
    P3Pixel p=p3_pixel_text(p3_field(even,0,0x20),p3_field(odd,0,0x20),p3_field(even,1,0x0a),drop&0xff,(drop>>8)&0xff,(drop>>16)&0xff);
    char t[16]; int n=sprintf(t,"%u %u %u\n",v[0],v[1],v[2]);
    unsigned char got[12]; memcpy(got,&p.w0,4); memcpy(got+4,&p.w1,4); memcpy(got+8,&p.w2,4);
    if ((int)p.bits!=8*n || memcmp(got,t,n)!=0){printf("pixel bad %u %u %u\n",v[0],v[1],v[2]);return 1;}
    for(int i=n;i<12;++i) if(got[i]){printf("pixel tail not zero %u %u %u\n",v[0],v[1],v[2]);return 1;}
  }
  // a thread's worth: n pixels appended at an arbitrary byte phase, as the kernel does
  for (long it=0;it<1000000;++it){
    unsigned char v[12];
    for(int k=0;k<12;++k){ int m=rand()%4; v[k]= m==0? rand()%10 : m==1? rand()%100 : rand()%256; }
    int n = (it%7==0)? 1+rand()%4 : 4;
    unsigned start=rand()%4;
    unsigned char buf[64]; memset(buf,0,64);
    uint32_t* b32=(uint32_t*)buf;
    uint32_t lo=0, fill8=8*start; unsigned widx=0;
    for(int px=0;px<n;++px){
      uint32_t w=v[3*px]|(v[3*px+1]<<8)|(v[3*px+2]<<16);
      uint32_t drop=0x10101010u-(p3_extra_digits4(w)<<3);
      P3Digits2 even=p3_digits2(w&0x00ff00ffu), odd=p3_digits2((w>>8)&0x00ff00ffu);
      P3Pixel p=p3_pixel_text(p3_field(even,0,0x20),p3_field(odd,0,0x20),p3_field(even,1,0x0a),drop&0xff,(drop>>8)&0xff,(drop>>16)&0xff);
      P3Append a=p3_append_pixel(lo,fill8,p);
      b32[widx]|=a.out0; if(a.nfull>=2) b32[widx+1]=a.out1; if(a.nfull==3) b32[widx+2]=a.out2;
      widx+=a.nfull; lo=a.lo; fill8=a.fill8;
    }
    if(fill8) b32[widx]|=lo;
    std::string want; for(int k=0;k<3*n;++k){ char t[8]; sprintf(t,"%u%c",v[k],(k%3==2)?'\n':' '); want+=t; }
    if(memcmp(buf+start,want.data(),want.size())!=0){printf("thread bad it=%ld n=%d start=%u\n",it,n,start);return 1;}
    for(unsigned i=start+want.size();i<64;++i) if(buf[i]){printf("trailing garbage\n");return 1;}
    for(unsigned i=0;i<start;++i) if(buf[i]){printf("leading garbage\n");return 1;}
  }
  printf("ok\n");
}
