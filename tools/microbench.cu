// Microbenchmarks that fix the design constants of the fused evidence kernel on B200 (sm_100a).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -cudart shared -o microbench microbench.cu
// Each test prints one line; numbers are per-SM rates derived from CUDA-event time and SM clock.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

__device__ __forceinline__ uint32_t lcg(uint32_t& s){ s = s*1664525u + 1013904223u; return s; }

// 1. shared-memory integer atomics at pseudo-random addresses
template<int NATOM>
__global__ void __launch_bounds__(1024,1) k_atoms(unsigned* out, int iters, int ncells){
  extern __shared__ unsigned sm[];
  for(int i=threadIdx.x;i<ncells*2;i+=blockDim.x) sm[i]=0;
  __syncthreads();
  uint32_t s = threadIdx.x*2654435761u + blockIdx.x*97u + 1u;
  for(int it=0; it<iters; ++it){
    uint32_t r = lcg(s);
    uint32_t c = (uint32_t)(((uint64_t)(r>>8) * (uint64_t)ncells) >> 24);
    atomicAdd(&sm[2*c], 1u);
    if (NATOM>=2) atomicAdd((int*)&sm[2*c+1], (int)(r&0xffff));
  }
  __syncthreads();
  unsigned acc=0; for(int i=threadIdx.x;i<ncells*2;i+=blockDim.x) acc+=sm[i];
  if(acc==0xdeadbeef) out[0]=acc;
}
// 2. global (L2) RED at pseudo-random addresses in a per-CTA 480 KB region
__global__ void __launch_bounds__(1024,1) k_redg(unsigned* g, int iters, int ncells){
  unsigned* mine = g + (size_t)blockIdx.x*ncells;
  uint32_t s = threadIdx.x*2654435761u + blockIdx.x*97u + 1u;
  for(int it=0; it<iters; ++it){
    uint32_t r = lcg(s);
    uint32_t c = (uint32_t)(((uint64_t)(r>>8) * (uint64_t)ncells) >> 24);
    atomicAdd(&mine[c], 1u);
  }
}
// 3. conversion / DFMA throughput (register-only)
__global__ void __launch_bounds__(1024,1) k_cvt(float* out, int iters, float seed){
  float a0=seed+threadIdx.x, a1=a0+1.f, a2=a0+2.f, a3=a0+3.f;
  for(int it=0; it<iters; ++it){
    #pragma unroll
    for(int u=0;u<4;++u){
      double d0=(double)a0, d1=(double)a1, d2=(double)a2, d3=(double)a3;
      d0+=1.0; d1+=1.0; d2+=1.0; d3+=1.0;
      a0=(float)d0; a1=(float)d1; a2=(float)d2; a3=(float)d3;
    }
  }
  if(a0+a1+a2+a3==-1.f) out[0]=a0;
}
__global__ void __launch_bounds__(1024,1) k_dfma(double* out, int iters, double seed){
  double a0=seed+threadIdx.x, a1=a0+1., a2=a0+2., a3=a0+3.;
  const double m=1.0000001, c=1e-9;
  for(int it=0; it<iters; ++it){
    #pragma unroll
    for(int u=0;u<8;++u){ a0=fma(a0,m,c); a1=fma(a1,m,c); a2=fma(a2,m,c); a3=fma(a3,m,c); }
  }
  if(a0+a1+a2+a3==-1.) out[0]=a0;
}
__global__ void __launch_bounds__(1024,1) k_ffma(float* out, int iters, float seed){
  float a0=seed+threadIdx.x, a1=a0+1.f, a2=a0+2.f, a3=a0+3.f;
  const float m=1.0000001f, c=1e-9f;
  for(int it=0; it<iters; ++it){
    #pragma unroll
    for(int u=0;u<8;++u){ a0=fmaf(a0,m,c); a1=fmaf(a1,m,c); a2=fmaf(a2,m,c); a3=fmaf(a3,m,c); }
  }
  if(a0+a1+a2+a3==-1.f) out[0]=a0;
}
// 4. TMA bulk (cp.async.bulk) streaming read, persistent CTAs, STAGES x TILE_BYTES ring
__device__ __forceinline__ void mbar_init(uint64_t* b, int n){ asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;"::"r"((uint32_t)__cvta_generic_to_shared(b)),"r"(n)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes){ asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"::"r"((uint32_t)__cvta_generic_to_shared(b)),"r"(bytes):"memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity){
  asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}"::"r"((uint32_t)__cvta_generic_to_shared(b)),"r"(parity):"memory"); }
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* b){
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
    ::"r"((uint32_t)__cvta_generic_to_shared(dst)),"l"(src),"r"(bytes),"r"((uint32_t)__cvta_generic_to_shared(b)):"memory"); }
template<int STAGES>
__global__ void __launch_bounds__(1024,1) k_tma(const float* __restrict__ in, size_t nbytes, int tile_bytes, float* out, int lds_per_pt){
  extern __shared__ __align__(128) unsigned char smraw[];
  __shared__ __align__(8) uint64_t full[STAGES];
  float* tiles = (float*)smraw;
  size_t ntiles = nbytes / tile_bytes;
  if(threadIdx.x==0){ for(int s=0;s<STAGES;++s) mbar_init(&full[s],1); asm volatile("fence.mbarrier_init.release.cluster;":::"memory"); }
  __syncthreads();
  size_t t0 = blockIdx.x;
  if(threadIdx.x==0){
    for(int s=0;s<STAGES-1;++s){ size_t t=t0+(size_t)s*gridDim.x; if(t<ntiles){ mbar_expect(&full[s],tile_bytes); bulk_g2s(smraw+(size_t)s*tile_bytes, (const char*)in+t*tile_bytes, tile_bytes, &full[s]); } }
  }
  float acc=0.f; int k=0;
  for(size_t t=t0; t<ntiles; t+=gridDim.x, ++k){
    int s = k%STAGES; uint32_t ph=(k/STAGES)&1;
    if(threadIdx.x==0){ size_t tn=t+(size_t)(STAGES-1)*gridDim.x; int sn=(k+STAGES-1)%STAGES; if(tn<ntiles){ mbar_expect(&full[sn],tile_bytes); bulk_g2s(smraw+(size_t)sn*tile_bytes,(const char*)in+tn*tile_bytes,tile_bytes,&full[sn]); } }
    mbar_wait(&full[s],ph);
    const float* tp = (const float*)(smraw+(size_t)s*tile_bytes);
    int nfl = tile_bytes/4; int npt = nfl/5;
    for(int p=threadIdx.x;p<npt;p+=blockDim.x){ for(int j=0;j<lds_per_pt;++j) acc += tp[p*5+j]; }
    __syncthreads();
  }
  if(acc==-1.2345f) out[0]=acc;
}
// plain vectorised LDG streaming read for comparison
__global__ void __launch_bounds__(1024,1) k_ldg(const float4* __restrict__ in, size_t n4, float* out){
  float acc=0.f;
  for(size_t i=(size_t)blockIdx.x*blockDim.x+threadIdx.x;i<n4;i+=(size_t)gridDim.x*blockDim.x){ float4 v=__ldg(in+i); acc+=v.x+v.y+v.z+v.w; }
  if(acc==-1.2345f) out[0]=acc;
}

int main(){
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,0));
  int nsm=p.multiProcessorCount; int clk_khz=0; cudaDeviceGetAttribute(&clk_khz,cudaDevAttrClockRate,0);
  printf("device %s sms %d clock_khz %d smem_optin %zu l2 %d\n",p.name,nsm,clk_khz,p.sharedMemPerBlockOptin,p.l2CacheSize);
  cudaEvent_t e0,e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  unsigned* dout; CK(cudaMalloc(&dout,1<<20));
  float ms;
  // --- atoms
  for(int ncells : {16384, 4096}){
   for(int nat=1; nat<=2; ++nat){
    size_t sh=(size_t)ncells*8; int iters=4096;
    if(nat==1){ CK(cudaFuncSetAttribute(k_atoms<1>,cudaFuncAttributeMaxDynamicSharedMemorySize,(int)sh)); k_atoms<1><<<nsm,1024,sh>>>(dout,16,ncells); CK(cudaDeviceSynchronize()); CK(cudaEventRecord(e0)); k_atoms<1><<<nsm,1024,sh>>>(dout,iters,ncells); CK(cudaEventRecord(e1)); }
    else      { CK(cudaFuncSetAttribute(k_atoms<2>,cudaFuncAttributeMaxDynamicSharedMemorySize,(int)sh)); k_atoms<2><<<nsm,1024,sh>>>(dout,16,ncells); CK(cudaDeviceSynchronize()); CK(cudaEventRecord(e0)); k_atoms<2><<<nsm,1024,sh>>>(dout,iters,ncells); CK(cudaEventRecord(e1)); }
    CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms,e0,e1));
    double ops=(double)iters*1024*nat; printf("ATOMS ncells=%d natom=%d: %.3f ms  %.2f Gatom/s/SM  (%.3f atom/ns/SM)\n",ncells,nat,ms,ops/ms/1e6,ops/ms/1e6);
   }
  }
  // --- redg
  { int ncells=120000; unsigned* g; CK(cudaMalloc(&g,(size_t)nsm*ncells*4)); CK(cudaMemset(g,0,(size_t)nsm*ncells*4));
    for(int iters : {64, 1024}){ k_redg<<<nsm,1024>>>(g,8,ncells); CK(cudaDeviceSynchronize()); CK(cudaEventRecord(e0)); k_redg<<<nsm,1024>>>(g,iters,ncells); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms,e0,e1));
      double ops=(double)iters*1024; printf("REDG iters=%d: %.3f ms  %.3f Gred/s/SM  chip %.1f Gred/s\n",iters,ms,ops/ms/1e6,ops*nsm/ms/1e6); }
    CK(cudaFree(g)); }
  // --- cvt / dfma / ffma
  { int iters=4096; k_cvt<<<nsm,1024>>>((float*)dout,8,1.f); CK(cudaDeviceSynchronize()); CK(cudaEventRecord(e0)); k_cvt<<<nsm,1024>>>((float*)dout,iters,1.f); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms,e0,e1));
    double ops=(double)iters*1024*4*4*2; printf("CVT f32<->f64 (+DADD): %.3f ms  %.2f Gcvt/s/SM  (plus %.2f Gdadd/s/SM)\n",ms,ops/ms/1e6,ops/2/ms/1e6); }
  { int iters=4096; k_dfma<<<nsm,1024>>>((double*)dout,8,1.); CK(cudaDeviceSynchronize()); CK(cudaEventRecord(e0)); k_dfma<<<nsm,1024>>>((double*)dout,iters,1.); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms,e0,e1));
    double ops=(double)iters*1024*32; printf("DFMA: %.3f ms  %.2f Gdfma/s/SM\n",ms,ops/ms/1e6); }
  { int iters=4096; k_ffma<<<nsm,1024>>>((float*)dout,8,1.f); CK(cudaDeviceSynchronize()); CK(cudaEventRecord(e0)); k_ffma<<<nsm,1024>>>((float*)dout,iters,1.f); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms,e0,e1));
    double ops=(double)iters*1024*32; printf("FFMA: %.3f ms  %.2f Gffma/s/SM\n",ms,ops/ms/1e6); }
  // --- streaming
  { size_t nbytes=(size_t)4<<30; float* in; CK(cudaMalloc(&in,nbytes)); CK(cudaMemset(in,0,nbytes));
    for(int rep=0;rep<2;++rep){ CK(cudaEventRecord(e0)); k_ldg<<<nsm*2,1024>>>((const float4*)in,nbytes/16,(float*)dout); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms,e0,e1)); }
    printf("LDG.128 stream: %.3f ms  %.1f GB/s\n",ms,nbytes/ms/1e6);
    for(int tile : {10240, 20480, 40960}){ for(int lds : {0, 4}){
      size_t sh=(size_t)4*tile; CK(cudaFuncSetAttribute(k_tma<4>,cudaFuncAttributeMaxDynamicSharedMemorySize,(int)sh));
      for(int rep=0;rep<2;++rep){ CK(cudaEventRecord(e0)); k_tma<4><<<nsm,1024,sh>>>(in,nbytes,tile,(float*)dout,lds); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms,e0,e1)); }
      printf("TMA bulk stream stages=4 tile=%d lds/pt=%d: %.3f ms  %.1f GB/s\n",tile,lds,ms,nbytes/ms/1e6); } }
    { int tile=20480; size_t sh=(size_t)8*tile; CK(cudaFuncSetAttribute(k_tma<8>,cudaFuncAttributeMaxDynamicSharedMemorySize,(int)sh));
      for(int rep=0;rep<2;++rep){ CK(cudaEventRecord(e0)); k_tma<8><<<nsm,1024,sh>>>(in,nbytes,tile,(float*)dout,4); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms,e0,e1)); }
      printf("TMA bulk stream stages=8 tile=%d lds/pt=4: %.3f ms  %.1f GB/s\n",tile,ms,nbytes/ms/1e6); }
    CK(cudaFree(in)); }
  CK(cudaGetLastError());
  printf("done\n");
  return 0;
}
