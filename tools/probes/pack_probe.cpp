#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include <emmintrin.h>
#include <immintrin.h>
typedef void (*fn_t)(const float*, long, long, float*);
static void v_nt(const float*src,long r0,long r1,float*dst){ const float*s=src+r0*5; float*d=dst+r0*3; for(long r=r0;r<r1;r++,s+=5,d+=3){ _mm_stream_si32((int*)d,((const int*)s)[0]);_mm_stream_si32((int*)d+1,((const int*)s)[1]);_mm_stream_si32((int*)d+2,((const int*)s)[2]);} _mm_sfence(); }
static void v_plain(const float*src,long r0,long r1,float*dst){ const float*s=src+r0*5; float*d=dst+r0*3; for(long r=r0;r<r1;r++,s+=5,d+=3){ d[0]=s[0];d[1]=s[1];d[2]=s[2];} }
// 4 rows (80 B in) -> 48 B out with 128-bit ops
static void v_sse(const float*src,long r0,long r1,float*dst){ const float*s=src+r0*5; float*d=dst+r0*3; long r=r0;
  for(;r+4<=r1;r+=4,s+=20,d+=12){ __m128 a=_mm_loadu_ps(s), b=_mm_loadu_ps(s+4), c=_mm_loadu_ps(s+8), e=_mm_loadu_ps(s+12), f=_mm_loadu_ps(s+16);
    // rows: r0: a0 a1 a2 | r1: b1 b2 b3 | r2: c2 c3 e0 | r3: e3 f0 f1
    __m128 o0=_mm_shuffle_ps(a, _mm_shuffle_ps(a,b,_MM_SHUFFLE(1,1,2,2)), _MM_SHUFFLE(2,0,1,0)); // a0 a1 a2 b1
    __m128 o1=_mm_shuffle_ps(_mm_shuffle_ps(b,b,_MM_SHUFFLE(3,3,3,2)), c, _MM_SHUFFLE(3,2,1,0)); // b2 b3 c2 c3
    __m128 o2=_mm_shuffle_ps(_mm_shuffle_ps(e,e,_MM_SHUFFLE(3,3,3,0)), f, _MM_SHUFFLE(1,0,1,0)); // e0 e3 f0 f1
    _mm_storeu_ps(d,o0); _mm_storeu_ps(d+4,o1); _mm_storeu_ps(d+8,o2); }
  for(;r<r1;r++,s+=5,d+=3){ d[0]=s[0];d[1]=s[1];d[2]=s[2]; } }
static void v_sse_nt(const float*src,long r0,long r1,float*dst){ const float*s=src+r0*5; float*d=dst+r0*3; long r=r0;
  for(;r<r1 && (((unsigned long)d)&15);r++,s+=5,d+=3){ d[0]=s[0];d[1]=s[1];d[2]=s[2]; }
  for(;r+4<=r1;r+=4,s+=20,d+=12){ __m128 a=_mm_loadu_ps(s), b=_mm_loadu_ps(s+4), c=_mm_loadu_ps(s+8), e=_mm_loadu_ps(s+12), f=_mm_loadu_ps(s+16);
    __m128 o0=_mm_shuffle_ps(a, _mm_shuffle_ps(a,b,_MM_SHUFFLE(1,1,2,2)), _MM_SHUFFLE(2,0,1,0));
    __m128 o1=_mm_shuffle_ps(_mm_shuffle_ps(b,b,_MM_SHUFFLE(3,3,3,2)), c, _MM_SHUFFLE(3,2,1,0));
    __m128 o2=_mm_shuffle_ps(_mm_shuffle_ps(e,e,_MM_SHUFFLE(3,3,3,0)), f, _MM_SHUFFLE(1,0,1,0));
    _mm_stream_ps(d,o0); _mm_stream_ps(d+4,o1); _mm_stream_ps(d+8,o2); }
  for(;r<r1;r++,s+=5,d+=3){ d[0]=s[0];d[1]=s[1];d[2]=s[2]; } _mm_sfence(); }
int main(int argc,char**argv){ long rows=41000000; float*src=(float*)aligned_alloc(64,rows*20); float*dst=(float*)aligned_alloc(64,rows*12); for(long i=0;i<rows*5;i++) src[i]=i; memset(dst,0,rows*12);
  struct {const char*n; fn_t f;} V[]={{"nt32",v_nt},{"plain",v_plain},{"sse",v_sse},{"sse_nt",v_sse_nt}};
  for(auto&v:V) for(int nt: {8,12,15,16}){ double best=1e9; for(int it=0;it<4;it++){ auto t0=std::chrono::steady_clock::now(); std::vector<std::thread> th; long per=(rows+nt-1)/nt; for(int t=0;t<nt;t++) th.emplace_back(v.f,src,t*per,std::min(rows,(t+1)*per),dst); for(auto&t:th)t.join(); double dt=std::chrono::duration<double>(std::chrono::steady_clock::now()-t0).count(); if(dt<best)best=dt;} 
    bool ok=true; for(long r=0;r<rows;r+=977) for(int k=0;k<3;k++) if(dst[r*3+k]!=src[r*5+k]) ok=false;
    printf("%s %d threads: %.1f ms  %.1f GB/s in  %s\n",v.n,nt,best*1e3,rows*20/best/1e9, ok?"ok":"BAD"); }
}
