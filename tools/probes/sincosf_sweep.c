// sincosf_sweep.c - exhaustive check of the restated glibc sinf/cosf (oracle/orb_oracle.c: orc_glibc_cosf) against the host libm.
// gcc -O2 -ffp-contract=off [-mfma -DUSE_FMA] sincosf_sweep.c -lm ; ./a.out 0x00000000 0x40c91000   (all floats in [0, 2 pi])
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
static const double SGN[4]={1.0,-1.0,-1.0,1.0};
static const double HPI_INV=0x1.45F306DC9C883p+23, HPI=0x1.921FB54442D18p0;
static const double C0=1.0, C1=-0x1.ffffffd0c621cp-2, C2=0x1.55553e1068f19p-5, C3=-0x1.6c087e89a359dp-10, C4=0x1.99343027bf8c3p-16;
static const double S1=-0x1.555545995a603p-3, S2=0x1.1107605230bc4p-7, S3=-0x1.994eb3774cf24p-13;
static inline uint32_t asuint(float f){uint32_t u;memcpy(&u,&f,4);return u;}
static inline uint32_t abstop12(float x){return (asuint(x)>>20)&0x7ff;}
#ifdef USE_FMA
#define MA(a,b,c) fma((a),(b),(c))
#else
#define MA(a,b,c) ((a)*(b)+(c))
#endif
static inline float poly(double x,double x2,int neg,int n){
  /* neg: use table[1] (negated cosine coefficients) */
  if((n&1)==0){
    double x3=x*x2; double s1=MA(x2,S3,S2); double x7=x3*x2; double s=MA(x3,S1,x); return (float)MA(x7,s1,s);
  } else {
    double sg=neg?-1.0:1.0;
    double x4=x2*x2; double c2=MA(x2,sg*C4,sg*C3); double c1=MA(x2,sg*C1,sg*C0); double x6=x4*x2; double c=MA(x4,sg*C2,c1); return (float)MA(x6,c2,c);
  }
}
static inline double reduce_fast(double x,int*np){ double r=x*HPI_INV; int n=((int32_t)r+0x800000)>>24; *np=n; return MA(-(double)n,HPI,x); }
float my_cosf(float y){ double x=y; int n;
  if(abstop12(y)<abstop12(0x1.921FB6p-1f)){ double x2=x*x; if(abstop12(y)<abstop12(0x1p-12f)) return 1.0f; return poly(x,x2,0,1);} 
  x=reduce_fast(x,&n); double s=SGN[n&3]; return poly(x*s,x*x,(n&2)!=0,n^1);
}
float my_sinf(float y){ double x=y; int n;
  if(abstop12(y)<abstop12(0x1.921FB6p-1f)){ double s=x*x; if(abstop12(y)<abstop12(0x1p-12f)) return y; return poly(x,s,0,0);} 
  x=reduce_fast(x,&n); double s=SGN[n&3]; return poly(x*s,x*x,(n&2)!=0,n);
}
int main(int argc,char**argv){
  uint32_t lo=strtoul(argv[1],0,0), hi=strtoul(argv[2],0,0); long bc=0,bs=0;
  for(uint32_t u=lo;u<hi;++u){ float f; memcpy(&f,&u,4); volatile float vf=f;
    float c=cosf(vf), s=sinf(vf);
    if(asuint(c)!=asuint(my_cosf(f))){ if(bc<3) printf("cos %a lib %a mine %a\n",f,c,my_cosf(f)); bc++; }
    if(asuint(s)!=asuint(my_sinf(f))){ if(bs<3) printf("sin %a lib %a mine %a\n",f,s,my_sinf(f)); bs++; }
  }
  printf("range %x..%x cos mismatches %ld sin mismatches %ld\n",lo,hi,bc,bs); return 0; }
