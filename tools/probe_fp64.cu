// FP64 peak probe for B200 (sm_100a): DMMA.8x8x4 issue rate, DFMA rate, cuBLAS Dgemm.
// Writes one JSON object to stdout. Build: see tools/Makefile target probe_fp64.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
#include <cublas_v2.h>

#define CK(x) do{cudaError_t e_=(x); if(e_!=cudaSuccess){fprintf(stderr,"CUDA %s @%d\n",cudaGetErrorString(e_),__LINE__); exit(1);} }while(0)

template<int ILP>
__global__ void __launch_bounds__(1024) dmma_loop(double* out, int iters, double a0, double b0) {
  double c[ILP][2];
#pragma unroll
  for (int i=0;i<ILP;++i){c[i][0]=threadIdx.x*1e-9; c[i][1]=i;}
  double a=a0+threadIdx.x*1e-12, b=b0;
  for (int it=0; it<iters; ++it) {
#pragma unroll
    for (int i=0;i<ILP;++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]),"+d"(c[i][1]) : "d"(a),"d"(b));
  }
  double s=0;
#pragma unroll
  for (int i=0;i<ILP;++i) s+=c[i][0]+c[i][1];
  if (s==123.456) out[0]=s;
}

template<int ILP>
__global__ void __launch_bounds__(1024) dfma_loop(double* out, int iters, double a0, double b0) {
  double c[ILP];
#pragma unroll
  for (int i=0;i<ILP;++i) c[i]=threadIdx.x*1e-9+i;
  double a=a0, b=b0;
  for (int it=0; it<iters; ++it) {
#pragma unroll
    for (int i=0;i<ILP;++i) c[i]=fma(c[i],a,b);
  }
  double s=0;
#pragma unroll
  for (int i=0;i<ILP;++i) s+=c[i];
  if (s==123.456) out[0]=s;
}

template<typename F> float time_ms(F f, int rep=5){
  cudaEvent_t e0,e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f(); CK(cudaDeviceSynchronize());
  float best=1e30f;
  for(int r=0;r<rep;++r){ CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms,e0,e1)); best=std::min(best,ms);} 
  return best;
}

int main(){
  int dev=0; CK(cudaSetDevice(dev)); cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,dev));
  int sms=p.multiProcessorCount;
  double* out; CK(cudaMalloc(&out,8));
  printf("{\"gpu\":\"%s\",\"sms\":%d,\"clock_khz\":%d,\n", p.name, sms, p.clockRate);
  printf("\"dmma\":[");
  bool first=true;
  int iters=20000;
  auto run_dmma=[&](int warps,int ilp){
    float ms=0;
    auto L=[&](auto K){ ms=time_ms([&]{K<<<sms,warps*32>>>(out,iters,1.0,1e-9);}); };
    if(ilp==1) L(dmma_loop<1>); else if(ilp==2) L(dmma_loop<2>); else if(ilp==4) L(dmma_loop<4>); else if(ilp==8) L(dmma_loop<8>); else L(dmma_loop<16>);
    double flops=2.0*256*(double)iters*ilp*warps*sms;
    printf("%s{\"warps\":%d,\"ilp\":%d,\"ms\":%.4f,\"tflops\":%.3f}", first?"":",", warps,ilp,ms,flops/ms*1e-9); first=false;
  };
  for(int w: {4,8,16,32}) for(int ilp: {1,2,4,8,16}) run_dmma(w,ilp);
  printf("],\n\"dfma\":["); first=true;
  auto run_dfma=[&](int warps,int ilp){
    float ms=0;
    auto L=[&](auto K){ ms=time_ms([&]{K<<<sms,warps*32>>>(out,iters*8,1.0000001,1e-9);}); };
    if(ilp==1) L(dfma_loop<1>); else if(ilp==2) L(dfma_loop<2>); else if(ilp==4) L(dfma_loop<4>); else if(ilp==8) L(dfma_loop<8>); else L(dfma_loop<16>);
    double flops=2.0*32*(double)iters*8*ilp*warps*sms;
    printf("%s{\"warps\":%d,\"ilp\":%d,\"ms\":%.4f,\"tflops\":%.3f}", first?"":",", warps,ilp,ms,flops/ms*1e-9); first=false;
  };
  for(int w: {8,16,32}) for(int ilp: {4,8,16}) run_dfma(w,ilp);
  printf("],\n");
  // cuBLAS Dgemm
  cublasHandle_t h; cublasCreate(&h);
  printf("\"cublas_dgemm\":["); first=true;
  for (int n : {2048, 4096, 8192}) {
    double *A,*B,*C; size_t bytes=(size_t)n*n*8;
    CK(cudaMalloc(&A,bytes)); CK(cudaMalloc(&B,bytes)); CK(cudaMalloc(&C,bytes));
    CK(cudaMemset(A,0,bytes)); CK(cudaMemset(B,0,bytes));
    double one=1, zero=0;
    for (int t=0;t<2;++t){
      cublasOperation_t ta = t? CUBLAS_OP_T: CUBLAS_OP_N;
      float ms=time_ms([&]{cublasDgemm(h,ta,CUBLAS_OP_N,n,n,n,&one,A,n,B,n,&zero,C,n);},10);
      double tf=2.0*n*(double)n*n/ms*1e-9;
      double sustained=0;
      if(n==8192){ // 4 s back-to-back
        cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        int reps=std::max(1,(int)(4000.0/ms));
        cudaEventRecord(e0); for(int r=0;r<reps;++r) cublasDgemm(h,ta,CUBLAS_OP_N,n,n,n,&one,A,n,B,n,&zero,C,n); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float tms; cudaEventElapsedTime(&tms,e0,e1); sustained=2.0*n*(double)n*n*reps/tms*1e-9;
      }
      printf("%s{\"n\":%d,\"opA\":\"%s\",\"ms\":%.4f,\"tflops\":%.3f,\"sustained_tflops\":%.3f}", first?"":",", n, t?"T":"N", ms, tf, sustained); first=false;
    }
    cudaFree(A);cudaFree(B);cudaFree(C);
  }
  printf("]}\n");
  return 0;
}
