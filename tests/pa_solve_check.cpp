// TEST: the product's row-per-lane 16x16 elimination (owg_pa_core.h pa_tile_solve16: reduce-based pivot search, row-index swaps, shared-memory
// rows) on the lane emulator against the sequential elimination of gen_power_amp.rs:9300-9345, on random systems with exact ties, zero
// columns and singular matrices.  Prints "bad=<mismatches> singular=<count>".
#include <cstdio>
#include <cstdint>
#include "pa_tile_emu.cpp"
static double A[16][16], B[16], X[16]; static bool OK[16]; static PaScratch SC;
void lane_solve(int lane){ EmuTile t{lane, 0}; for(int j=0;j<16;j++) SC.a[lane][j]=A[lane][j]; SC.a[lane][16]=B[lane]; OK[lane]=pa_tile_solve16(t,SC); 
  g_emu->done[lane]=true; for(int i=1;i<=L;i++){int nxt=(lane+i)%L; if(!g_emu->done[nxt]){g_emu->cur=nxt; setcontext(&g_emu->ctx[nxt]);}} setcontext(&g_emu->main_ctx);}
bool seq(double a[16][16], double b[16]){
  for(int col=0;col<16;col++){int mr=col; double mv=fabs(a[col][col]); for(int r=col+1;r<16;r++) if(fabs(a[r][col])>mv){mv=fabs(a[r][col]);mr=r;}
    if(mv<1e-15) return false; if(mr!=col){for(int j=0;j<16;j++){double t=a[col][j];a[col][j]=a[mr][j];a[mr][j]=t;} double t=b[col];b[col]=b[mr];b[mr]=t;}
    double p=a[col][col]; for(int r=col+1;r<16;r++){double f=a[r][col]/p; for(int j=col+1;j<16;j++) a[r][j]-=f*a[col][j]; b[r]-=f*b[col];}}
  for(int i=15;i>=0;i--){double s=b[i]; for(int j=i+1;j<16;j++) s-=a[i][j]*b[j]; if(fabs(a[i][i])<1e-15) return false; b[i]=s/a[i][i];} return true;}
int main(){ srand(1); int bad=0, sing=0;
 for(int trial=0;trial<400;trial++){
  for(int i=0;i<16;i++){for(int j=0;j<16;j++) A[i][j]=(rand()/(double)RAND_MAX-0.5)*(trial%3==0?1.0:0.1)+(i==j&&trial%2?1.0:0.0); B[i]=rand()/(double)RAND_MAX-0.5;}
  if(trial%7==3){ for(int j=0;j<16;j++) A[5][j]=A[9][j]; }
  if(trial%11==5){ A[3][0]=-A[0][0]; A[7][0]=A[0][0]; }   // exact ties in column 0
  if(trial%13==6){ for(int i=0;i<16;i++) A[i][2]=0.0; }     // zero column
  double a2[16][16],b2[16]; memcpy(a2,A,sizeof(A)); memcpy(b2,B,sizeof(B)); bool ok2=seq(a2,b2); if(!ok2) sing++;
  Emu emu; g_emu=&emu; emu.nl=L; for(int i=0;i<L;i++){emu.done[i]=false; emu.stacks[i].resize(1<<18); getcontext(&emu.ctx[i]); emu.ctx[i].uc_stack.ss_sp=emu.stacks[i].data(); emu.ctx[i].uc_stack.ss_size=emu.stacks[i].size(); emu.ctx[i].uc_link=nullptr; makecontext(&emu.ctx[i],(void(*)())lane_solve,1,i);} emu.cur=0; swapcontext(&emu.main_ctx,&emu.ctx[0]);
  bool mism=false; for(int l=0;l<16;l++) if(OK[l]!=ok2) mism=true; if(ok2) for(int j=0;j<16;j++) if(SC.x[j]!=b2[j]) mism=true;
  if(mism){bad++; if(bad<5) printf("trial %d mismatch ok=%d/%d x0 %.17g vs %.17g\n",trial,OK[0],ok2,SC.x[0],b2[0]);}
 }
 printf("bad=%d singular=%d\n",bad,sing); }
