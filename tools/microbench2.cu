// Conversion-pipe throughput on B200: F2F f32<->f64, F2I trunc / rint, I2F, and a magic-number rint alternative.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)
template<int MODE>
__global__ void __launch_bounds__(1024,1) k(float* out, int iters, float seed, double dm){
  float a0=seed+threadIdx.x*0.37f, a1=a0+1.3f, a2=a0+2.7f, a3=a0+3.1f;
  int acc=0;
  for(int it=0; it<iters; ++it){
    #pragma unroll
    for(int u=0;u<8;++u){
      if(MODE==0){ // f32->f64, dmul by non-f32 constant, f64->f32 : 2 cvt + 1 DMUL per value
        a0=(float)((double)a0*dm); a1=(float)((double)a1*dm); a2=(float)((double)a2*dm); a3=(float)((double)a3*dm);
      } else if(MODE==1){ // F2I trunc + dependent FADD
        acc+=(int)a0; acc+=(int)a1; acc+=(int)a2; acc+=(int)a3; a0+=0.37f; a1+=0.37f; a2+=0.37f; a3+=0.37f;
      } else if(MODE==2){ // F2I rint
        acc+=__float2int_rn(a0); acc+=__float2int_rn(a1); acc+=__float2int_rn(a2); acc+=__float2int_rn(a3); a0+=0.37f; a1+=0.37f; a2+=0.37f; a3+=0.37f;
      } else if(MODE==3){ // magic-number rint (FADD + LOP)
        acc+=__float_as_int(a0+8388608.0f)&0x7fffff; acc+=__float_as_int(a1+8388608.0f)&0x7fffff; acc+=__float_as_int(a2+8388608.0f)&0x7fffff; acc+=__float_as_int(a3+8388608.0f)&0x7fffff; a0+=0.37f; a1+=0.37f; a2+=0.37f; a3+=0.37f;
      } else if(MODE==4){ // I2F
        a0+=(float)(acc+u); a1+=(float)(acc+u+1); a2+=(float)(acc+u+2); a3+=(float)(acc+u+3); acc+=3;
      } else if(MODE==5){ // baseline: FADD + IADD only
        acc+=__float_as_int(a0); acc+=__float_as_int(a1); acc+=__float_as_int(a2); acc+=__float_as_int(a3); a0+=0.37f; a1+=0.37f; a2+=0.37f; a3+=0.37f;
      } else if(MODE==6){ // IEEE fdiv by 100
        a0=__fdiv_rn(a0,100.0f)+1e3f; a1=__fdiv_rn(a1,100.0f)+1e3f; a2=__fdiv_rn(a2,100.0f)+1e3f; a3=__fdiv_rn(a3,100.0f)+1e3f;
      } else if(MODE==7){ // f32 sqrt_rn
        a0=__fsqrt_rn(a0)+1e3f; a1=__fsqrt_rn(a1)+1e3f; a2=__fsqrt_rn(a2)+1e3f; a3=__fsqrt_rn(a3)+1e3f;
      }
    }
  }
  if(a0+a1+a2+a3+(float)acc==-1.f) out[0]=a0;
}
template<int MODE> void run(const char* name, int ops_per_val, float* dout){
  cudaEvent_t e0,e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); int iters=2048; float ms;
  k<MODE><<<148,1024>>>(dout,8,1.5f,1.0000000001); CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0)); k<MODE><<<148,1024>>>(dout,iters,1.5f,1.0000000001); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms,e0,e1));
  double vals=(double)iters*8*4*1024; printf("%-34s %.3f ms  %.1f Gval/s/SM  (%d cvt-class ops per val -> %.1f Gop/s/SM)\n",name,ms,vals/ms/1e6,ops_per_val,vals*ops_per_val/ms/1e6);
}
int main(){ float* dout; CK(cudaMalloc(&dout,1024));
  run<0>("F2F.64.32 + DMUL + F2F.32.64",2,dout); run<1>("F2I.TRUNC",1,dout); run<2>("F2I.RN",1,dout); run<3>("magic rint (FADD+LOP)",1,dout);
  run<4>("I2F",1,dout); run<5>("baseline FADD+IADD",1,dout); run<6>("__fdiv_rn(x,100)",1,dout); run<7>("__fsqrt_rn",1,dout);
  CK(cudaGetLastError()); printf("done\n"); return 0; }
