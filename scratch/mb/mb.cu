#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
template<int MODE>
__global__ void k(uint32_t* out, int iters, uint32_t seed){
  __shared__ uint32_t sm[2048];
  for(int i=threadIdx.x;i<2048;i+=blockDim.x) sm[i]=0;
  __syncthreads();
  uint32_t v = (threadIdx.x*2654435761u) ^ seed; uint32_t acc=0;
  int lane=threadIdx.x&31;
  for(int i=0;i<iters;++i){
    v = v*1664525u+1013904223u;
    uint32_t key = (v>>20)&31;  // 32 possible values -> several lanes share
    if(MODE==0){ acc += __match_any_sync(0xffffffffu,key); }
    else if(MODE==1){ acc += __shfl_sync(0xffffffffu,key,(lane+1)&31); }
    else if(MODE==2){ atomicOr(&sm[(threadIdx.x>>5)*64+key], 1u<<lane); }
    else if(MODE==3){ acc += __ballot_sync(0xffffffffu,key&1); }
    else if(MODE==4){ sm[(threadIdx.x>>5)*64+lane]=key; acc+=sm[(threadIdx.x>>5)*64+((lane+1)&31)]; }
    else if(MODE==5){ acc += __reduce_or_sync(0xffffffffu,key); }
    else if(MODE==6){ // emulate cheap all-pairs compare via 16 shuffles
      #pragma unroll
      for(int j=1;j<16;++j){ uint32_t o=__shfl_sync(0xffffffffu,key,(lane+j)&31); acc |= ((o-key+3u)<7u)<<j; }
    }
  }
  if(MODE==2){ __syncthreads(); acc=sm[threadIdx.x]; }
  out[blockIdx.x*blockDim.x+threadIdx.x]=acc+v;
}
template<int MODE> void run(const char* name){
  uint32_t* d; cudaMalloc(&d, 148*8*256*4);
  int iters=2000; cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b);
  k<MODE><<<148*8,256>>>(d,iters,1); cudaDeviceSynchronize();
  cudaEventRecord(a); k<MODE><<<148*8,256>>>(d,iters,2); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms,a,b);
  // per SM: 8 CTAs*8 warps = 64 warps, iters each
  double cyc = ms*1e-3*1.965e9; double per = cyc/(64.0*iters);
  printf("%-28s %.3f ms  -> %.2f SM-cycles per warp-instruction-group\n",name,ms,per);
  cudaFree(d);
}
int main(){
  run<1>("shfl (baseline loop)"); run<0>("match_any"); run<2>("atomicOr smem (spread)"); run<3>("ballot"); run<4>("sts+lds"); run<5>("reduce_or"); run<6>("15 shfl+cmp");
  return 0;
}
