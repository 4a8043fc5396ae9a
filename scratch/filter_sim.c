// Filter-configuration simulator: greedy search (W=C=128) on the C2 graph, counts rows fetched for several filters.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <float.h>
#define W 128
typedef struct { float d; uint32_t id; int vis; } E;
static float l2(const float* a, const float* b, int dim){ float s=0; for(int i=0;i<dim;i++){float c=a[i]-b[i]; s+=c*c;} return s; }
// filter: sets x ways of 16-bit tags, LRU by position (insert at front)
typedef struct { int sets, ways; uint32_t* tags; int promote; } F;
static int fhit(F* f, uint32_t id){
  uint32_t set = id % f->sets, tag = id / f->sets; uint32_t* t = f->tags + (size_t)set*f->ways;
  for(int w=0; w<f->ways; w++) if(t[w]==tag){ if(f->promote){ for(int k=w;k>0;k--) t[k]=t[k-1]; t[0]=tag;} return 1; }
  for(int k=f->ways-1;k>0;k--) t[k]=t[k-1]; t[0]=tag; return 0; }
int main(int argc, char** argv){
  int n=1000000, dim=96, R=64, nq=atoi(argv[4]);
  float* x = malloc((size_t)n*dim*4); uint32_t* g = malloc((size_t)n*(R+1)*4); float* q = malloc((size_t)10000*dim*4);
  FILE* f=fopen(argv[1],"rb"); fread(x,4,(size_t)n*dim,f); fclose(f);
  f=fopen(argv[2],"rb"); fread(g,4,(size_t)n*(R+1),f); fclose(f);
  f=fopen(argv[3],"rb"); fread(q,4,(size_t)10000*dim,f); fclose(f);
  uint32_t ep = atoi(argv[5]);
  int cfgs[][3] = {{2048,2,0},{256,8,0},{256,8,1},{512,4,0},{128,8,0},{128,16,0},{192,8,0},{224,8,0},{512,8,0}};
  int ncfg = sizeof(cfgs)/sizeof(cfgs[0]);
  double fetched[32]={0}; double distinct=0, evals=0, hops=0; double survs=0, groups=0, nosurv=0; double candhist[70]={0};
  uint8_t* seen = calloc(n,1);
  for(int c=-1;c<ncfg;c++){
    F fl; if(c>=0){ fl.sets=cfgs[c][0]; fl.ways=cfgs[c][1]; fl.promote=cfgs[c][2]; fl.tags=malloc((size_t)fl.sets*fl.ways*4);} 
    for(int qi=0; qi<nq; qi++){
      const float* qv=q+(size_t)qi*dim; E buf[W+1]; int size=0;
      if(c>=0) memset(fl.tags,0xff,(size_t)fl.sets*fl.ways*4);
      uint32_t touched[20000]; int nt=0;
      buf[0].d=l2(qv,x+(size_t)ep*dim,dim); buf[0].id=ep; buf[0].vis=0; size=1;
      for(;;){
        int pos=-1; for(int i=0;i<size;i++) if(!buf[i].vis){pos=i;break;} if(pos<0) break;
        buf[pos].vis=1; uint32_t node=buf[pos].id; const uint32_t* row=g+(size_t)node*(R+1); int deg=row[0];
        if(c<0) hops++;
        int ncand=0, nsurv=0;
        float back0 = buf[size-1].d; int full0 = size==W;
        for(int j=0;j<deg;j++){ uint32_t nb=row[1+j];
          if(c<0){ evals++; if(!seen[nb]){seen[nb]=1; touched[nt++]=nb; distinct++;} }
          if(c>=0){ if(fhit(&fl,nb)) continue; fetched[c]++; ncand++; }
          float d=l2(qv,x+(size_t)nb*dim,dim);
          if(size==W && buf[size-1].d < d) continue;
          if(c==0 && !(full0 && back0<d)) nsurv++;
          int ip=0; while(ip<size && buf[ip].d<=d) ip++;
          int dup=0; for(int k=ip-1;k>=0 && buf[k].d==d;k--) if(buf[k].id==nb){dup=1;break;}
          if(dup) continue;
          int ns = size<W? size+1: W; for(int k=ns-1;k>ip;k--) buf[k]=buf[k-1]; if(ip<ns){buf[ip].d=d;buf[ip].id=nb;buf[ip].vis=0;} size=ns;
        }
        if(c==0){ candhist[ncand>64?64:ncand]++; if(ncand){groups++; survs+=nsurv; if(!nsurv) nosurv++;} }
      }
      if(c<0) for(int i=0;i<nt;i++) seen[touched[i]]=0;
    }
    if(c<0) printf("hops %.1f evals %.1f distinct %.1f per query\n", hops/nq, evals/nq, distinct/nq);
    else printf("sets %5d ways %d promote %d bytes %6d: fetched %.1f\n", cfgs[c][0],cfgs[c][1],cfgs[c][2],cfgs[c][0]*cfgs[c][1]*2, fetched[c]/nq);
    if(c==0){ printf("hops with cand: %.1f/query, avg surv %.2f, frac no-surv %.3f\n", groups/nq, survs/groups, nosurv/groups);
      double cum=0, tot=0; for(int i=0;i<=64;i++) tot+=candhist[i];
      printf("ncand cdf: "); for(int i=0;i<=64;i++){cum+=candhist[i]; if(i==0||i==4||i==8||i==12||i==16||i==24||i==32||i==48||i==64) printf("<=%d:%.3f ",i,cum/tot);} printf("\n"); }
  }
  return 0; }
