// Issue-rate probe for the integer instructions the BS kernels lean on (sm_100a): how many warp
// instructions per cycle and SM sub-partition each sustains alone and mixed. Build + run:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/pipes tools/ubench/pipes.cu && gpurun_out/pipes
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 4096, UNROLL = 8;

template <int MODE>
__global__ void probe(unsigned *out, unsigned seed, long long *cycles) {
	unsigned x[UNROLL];
#pragma unroll
	for (int i = 0; i < UNROLL; i++) x[i] = seed + threadIdx.x * 7 + i;
	const unsigned c = seed | 0x80000001u;
	long long t0 = clock64();
	for (int it = 0; it < ITERS; it++) {
#pragma unroll
		for (int i = 0; i < UNROLL; i++) {
			if (MODE == 0) asm volatile("mad.lo.u32 %0, %0, %1, %0;" : "+r"(x[i]) : "r"(c));              // IMAD
			if (MODE == 1) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(c));                  // IMAD.HI.U32
			if (MODE == 2) asm volatile("shf.r.wrap.b32 %0, %0, %1, 7;" : "+r"(x[i]) : "r"(c));           // SHF
			if (MODE == 3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(c), "r"(seed)); // LOP3
			if (MODE == 4) {                                                                               // IMAD.HI + SHF
				if (i & 1) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(c));
				else asm volatile("shf.r.wrap.b32 %0, %0, %1, 7;" : "+r"(x[i]) : "r"(c));
			}
			if (MODE == 5) {                                                                               // IMAD + SHF
				if (i & 1) asm volatile("mad.lo.u32 %0, %0, %1, %0;" : "+r"(x[i]) : "r"(c));
				else asm volatile("shf.r.wrap.b32 %0, %0, %1, 7;" : "+r"(x[i]) : "r"(c));
			}
			if (MODE == 6) {                                                                               // IMAD + IMAD.HI
				if (i & 1) asm volatile("mad.lo.u32 %0, %0, %1, %0;" : "+r"(x[i]) : "r"(c));
				else asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(c));
			}
			if (MODE == 7) asm volatile("mad.hi.u32 %0, %0, %1, %0;" : "+r"(x[i]) : "r"(c));              // IMAD.HI with addend
			if (MODE == 8) asm volatile("min.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(c));                     // VIMNMX
			if (MODE == 9) asm volatile("prmt.b32 %0, %0, %1, 0x3021;" : "+r"(x[i]) : "r"(c));            // PRMT
		}
	}
	long long t1 = clock64();
	unsigned acc = 0;
#pragma unroll
	for (int i = 0; i < UNROLL; i++) acc ^= x[i];
	out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
	if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int MODE>
void run(const char *name, unsigned *out, long long *cyc) {
	// 16 warps per SM sub-partition worth of one CTA per SM: 512 threads
	probe<MODE><<<148, 512>>>(out, 12345u, cyc);
	cudaDeviceSynchronize();
	probe<MODE><<<148, 512>>>(out, 12345u, cyc);
	cudaDeviceSynchronize();
	long long h = 0;
	cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
	// warp instructions per sub-partition = 4 warps x ITERS x UNROLL
	printf("%-22s %8lld cycles  %.3f warp-instr / cycle / sub-partition\n", name, h, 4.0 * ITERS * UNROLL / (double)h);
}

int main() {
	unsigned *out; long long *cyc;
	cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 8);
	run<0>("IMAD", out, cyc);
	run<1>("IMAD.HI.U32", out, cyc);
	run<7>("IMAD.HI.U32 + addend", out, cyc);
	run<2>("SHF", out, cyc);
	run<3>("LOP3", out, cyc);
	run<8>("VIMNMX", out, cyc);
	run<9>("PRMT", out, cyc);
	run<4>("IMAD.HI | SHF", out, cyc);
	run<5>("IMAD | SHF", out, cyc);
	run<6>("IMAD | IMAD.HI", out, cyc);
	return 0;
}
