// host-side check of lz4_tpb.cuh against the oracle (valid + corrupted streams, all alignments)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <random>
#include "lz4_tpb.cuh"
extern "C" {
int64_t orc_lz4_compress(const uint8_t*,int64_t,uint8_t*,int64_t);
int64_t orc_lz4_decompress(const uint8_t*,int64_t,uint8_t*,int64_t,int64_t*);
}
int main(int argc,char**argv){
  FILE*f=fopen(argv[1],"rb"); fseek(f,0,SEEK_END); long sz=ftell(f); fseek(f,0,SEEK_SET); std::vector<uint8_t> d(sz); if(fread(d.data(),1,sz,f)!=(size_t)sz) return 1;
  std::mt19937 rng(42); long nok=0,nfb=0,nbad=0,nvalid_fb=0,ntests=0;
  std::vector<int> sizes={65536,4096,1000,131072,100,33,17,5,0};
  std::vector<uint8_t> cbuf(300000), ref(300000), obuf(300000+256), ibuf(300000+64);
  for(int bs: sizes){ int stride = bs>=4096? bs*3 : bs*97+1; if(bs==0){stride=sz;}
   for(long o=0;o+bs<=sz;o+=stride){
    int n=bs; int cl=(int)orc_lz4_compress(d.data()+o,n,cbuf.data(),cbuf.size());
    for(int variant=0;variant<6;variant++){
      std::vector<uint8_t> s(cbuf.begin(),cbuf.begin()+cl); int cap=n;
      if(variant==1) cap=n+1021;
      if(variant==2 && cl>4) s.resize(rng()%cl+1);
      if(variant==3) for(int k=0;k<1+(int)(rng()%3);k++) s[rng()%s.size()]^=1<<(rng()%8);
      if(variant==4 && n>0) cap=rng()%n;
      if(variant==5) s[rng()%std::min<size_t>(s.size(),64)]=rng();
      int ia=rng()%16, oa=rng()%16;
      memcpy(ibuf.data()+ia,s.data(),s.size());
      memset(obuf.data(),0xA5,obuf.size());
      uint32_t olen=0; int r=lz4tpb::decode_block(ibuf.data()+ia,(uint32_t)s.size(),obuf.data()+64+oa,(uint32_t)cap,&olen);
      int64_t eo=0; int64_t rr=orc_lz4_decompress(s.data(),s.size(),ref.data(),cap,&eo);
      ntests++;
      // guards
      for(int k=0;k<64+oa;k++) if(obuf[k]!=0xA5){printf("GUARD-before bs=%d o=%ld var=%d k=%d\n",bs,o,variant,k);return 2;}
      for(int k=0;k<128;k++) if(obuf[64+oa+cap+k]!=0xA5){printf("GUARD-after bs=%d o=%ld var=%d k=%d r=%d cap=%d olen=%u\n",bs,o,variant,k,r,cap,olen);return 2;}
      if(r==lz4tpb::kOk){ nok++; if(rr<0||rr!=(int64_t)olen||memcmp(obuf.data()+64+oa,ref.data(),olen)){printf("MISMATCH bs=%d o=%ld var=%d rr=%ld olen=%u\n",bs,o,variant,(long)rr,olen);return 3;} }
      else { nfb++; if(rr>=0){ nvalid_fb++; } else nbad++; }
    }
   }
  }
  printf("tests %ld ok %ld fallback %ld (valid streams falling back %ld, invalid %ld)\n",ntests,nok,nfb,nvalid_fb,nbad); return 0; }
