#include <cstdio>
#include <cstdint>
#include <random>
typedef unsigned long long u64;
static u64 brev(u64 x){u64 r=0;for(int i=0;i<64;i++) if(x>>i&1) r|=1ull<<(63-i);return r;}
static u64 fill_up(u64 f,u64 p){const u64 o=~p|f;return f|((o^(o-(f<<1)))&p);}
static u64 fill_down(u64 f,u64 p){return brev(fill_up(brev(f),brev(p)));}
static u64 low_run_incl(u64 a){return a^(a+1ull);}
static u64 low_run(u64 a){return a&~(a+1ull);}
static u64 high_run_incl(u64 a){u64 m=~a;return m?(~0ull<<(63-__builtin_clzll(m))):~0ull;}
static u64 high_run(u64 a){u64 m=~a;return m?~(~0ull>>__builtin_clzll(m)):~0ull;}
int main(){
  std::mt19937_64 g(1);
  long bad=0;
  for(long it=0;it<4000000;it++){
    u64 aU=g(),aD=g(),f=g()&g()&g(),b=g()&g()&g();
    int mode=it%6;
    if(mode==1){aU|=g()|g();aD|=g()|g();}
    if(mode==2){aU=~0ull;aD=~0ull;}
    if(mode==3){int k=g()%64;aU=~0ull>>k;aD=~0ull<<k;}
    if(mode==4){int k=g()%64;aU=~(1ull<<k);aD=~(1ull<<k);}
    if(mode==5){int k=g()%65;aU=k==64?~0ull:((1ull<<k)-1);aD=k==0?0:(~0ull<<(64-k));}
    // up pass
    u64 f1=fill_up(f,aU<<1),b1=fill_up(b,aD);
    if(fill_up(f1|1ull,aU<<1)!=(f1|low_run_incl(aU))) bad++;
    if(fill_up(b1|(aD&1ull),aD)!=(b1|low_run(aD))) bad++;
    // down pass
    u64 f2=fill_down(f,aD>>1),b2=fill_down(b,aU);
    if(fill_down(f2|(1ull<<63),aD>>1)!=(f2|high_run_incl(aD))) bad++;
    if(fill_down(b2|(aU&(1ull<<63)),aU)!=(b2|high_run(aU))) bad++;
    // frontier tests: empty iff the pass changes nothing inside the word
    u64 fa=f&aU; bool fr=(((fa<<1)&~f)|((b<<1)&aD&~b))!=0; if(fr!=((f1!=f)||(b1!=b))) bad++;
    u64 fd=f&aD; bool fr2=(((fd>>1)&~f)|((b>>1)&aU&~b))!=0; if(fr2!=((f2!=f)||(b2!=b))) bad++;
  }
  printf("bad=%ld\n",bad); return bad!=0;
}
