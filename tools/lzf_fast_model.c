// lzf_fast_model.c -- development tool: what an LZF "fast mode" built like the LZ4 one (window-parallel parse, all positions
// inserted, 3-byte keys, greedy, exact extension) costs in LZF format against the restated reference lzf_compress.
// Result: the table size decides -- 8 KiB chunks: 12 bits +1.5 % / worst file +2.5 %, 13 bits +0.6 % / +1.2 %,
// 14 bits +0.2 % / +0.5 %; 4 KiB: 13 bits +0.3 % / +0.7 %; 64 KiB: 14 bits -0.4 % / +0.8 % -- so the kernels run 13 bits up to
// 4 KiB and 14 above (cw_lzff_hashlog).  A cap of 32 matches per 128 positions (one lane per match) costs another 0.5-1.4 %,
// which is why an LZF group has three windows.
// gcc -O2 tools/lzf_fast_model.c -Ioracle -Loracle -loracle -Wl,-rpath,$PWD/oracle ; a.out <chunk> <bits> <per> <back> files...
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "cw_oracle.h"
static uint32_t rd24(const uint8_t *s, uint32_t p){return s[p]|(s[p+1]<<8)|(s[p+2]<<16);}
static int HL=12, PER=2, BACK=1;
static uint32_t lits_cost(uint32_t lit){ return lit + (lit+31)/32; }
static uint32_t match_cost(uint32_t len){ uint32_t c=0; while(len>264){ uint32_t take = len-264<3 ? 262 : 264; c+=3; len-=take; } return c + (len-2<7?2:3); }
static uint32_t model(const uint8_t*s,uint32_t n){
  static uint16_t table[1<<16];
  memset(table,0,sizeof(uint16_t)<<HL);
  uint32_t size=0,anchor=0,ip=1;
  if(n>=8){
    const uint32_t last=n-7;   // last position a match may start at: the 3 + 4 bytes it looks at lie inside the chunk
    for(uint32_t base=0;base<=last;base+=32){
      uint32_t cand[32],v[32],h[32]; int ok[32];
      for(int l=0;l<32;l++){uint32_t p=base+l; ok[l]=0; if(p>last)continue; v[l]=rd24(s,p); h[l]=(v[l]*2654435761u)>>(32-HL); cand[l]=table[h[l]];}
      for(int l=31;l>=0;l--) if(base+l<=last&&base+l!=0) table[h[l]]=(uint16_t)(base+l);
      for(int l=0;l<32;l++){uint32_t p=base+l; if(p>last||p==0)continue; uint32_t w=table[h[l]]; if(w<p)cand[l]=w;
        if(PER&&p>=(uint32_t)PER&&rd24(s,p-PER)==v[l])cand[l]=p-PER;
        ok[l]= rd24(s,cand[l])==v[l] && p-cand[l]<=8192; }
      for(int l= ip>base?(int)(ip-base):0; l<32; l++){
        if(!ok[l])continue;
        uint32_t cur=base+l,c=cand[l],e=cur+3,ce=c+3;
        while(e<n&&s[e]==s[ce])e++,ce++;
        if(BACK) while(cur>anchor&&c>0&&s[cur-1]==s[c-1]&&cur-c<=8192){cur--;c--;}
        size+=lits_cost(cur-anchor)+match_cost(e-cur);
        anchor=ip=e; l=(int)(e-base)-1;
      }
    }
  }
  size+=lits_cost(n-anchor);
  return size>=n? n : size;
}
int main(int argc,char**argv){
  uint32_t chunk=atoi(argv[1]); HL=atoi(argv[2]); PER=atoi(argv[3]); BACK=atoi(argv[4]);
  uint8_t*ref=malloc(2*chunk+64);
  unsigned long long T=0,R=0;double worst=0;
  for(int a=5;a<argc;a++){FILE*f=fopen(argv[a],"rb");if(!f)continue;fseek(f,0,SEEK_END);long sz=ftell(f);fseek(f,0,SEEK_SET);
    uint8_t*d=malloc(sz+1);if(fread(d,1,sz,f)!=(size_t)sz)return 3;fclose(f);
    unsigned long long t=0,r=0;
    for(long i=0;i+chunk<=sz;i+=chunk){t+=model(d+i,chunk); int l=cwo_lzf_compress(d+i,chunk,ref,chunk-1); r+= l? (unsigned)l: chunk;}
    if(r){double q=(double)t/r; if(q>worst)worst=q; if(getenv("V"))printf("%-30s %.4f\n",strrchr(argv[a],'/')+1,q);}
    T+=t;R+=r;free(d);}
  printf("chunk %u HL %d PER %d BACK %d: total %.4f worst %.4f\n",chunk,HL,PER,BACK,(double)T/R,worst);
  return 0;}
