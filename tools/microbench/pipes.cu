// Issue-rate microbenchmark for the integer instructions the scramble kernel is made of (sm_100a).
// Each warp runs ITER x 8 independent instructions of one kind; reported: warp-instructions per clock per SM sub-partition
// with W warps resident per sub-partition.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pipes pipes.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITER 2048

template <int OP>
__device__ __forceinline__ void op8(uint32_t (&r)[8], uint32_t s, uint32_t t) {
#pragma unroll
	for (int i = 0; i < 8; ++i) {
		if (OP == 0) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(t), "r"(s));            // PRMT, register selector
		if (OP == 1) asm volatile("prmt.b32 %0, %0, %1, 0x3210;" : "+r"(r[i]) : "r"(t));                 // PRMT, immediate selector
		if (OP == 2) asm volatile("lop3.b32 %0, %0, %1, %2, 0x78;" : "+r"(r[i]) : "r"(s), "r"(t));       // LOP3 3 regs
		if (OP == 3) asm volatile("lop3.b32 %0, %0, 0x03030303, %1, 0x78;" : "+r"(r[i]) : "r"(t));       // LOP3 imm
		if (OP == 4) asm volatile("add.u32 %0, %0, %1;" : "+r"(r[i]) : "r"(s));                          // IADD (compiler's choice of pipe)
		if (OP == 5) asm volatile("mad.lo.u32 %0, %0, %2, %1;" : "+r"(r[i]) : "r"(s), "r"(t));                  // IMAD imm
		if (OP == 6) asm volatile("mul.hi.u32 %0, %0, 65536;" : "+r"(r[i]));                             // IMAD.HI
		if (OP == 7) asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(s), "r"(t));         // IDP.4A
		if (OP == 8) asm volatile("shr.u32 %0, %0, 3;" : "+r"(r[i]));                                    // SHF
		if (OP == 9) {                                                                                   // PRMT + IMAD.HI alternating
			if (i & 1) asm volatile("mul.hi.u32 %0, %0, 65536;" : "+r"(r[i]));
			else asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(t), "r"(s));
		}
		if (OP == 10) {                                                                                  // PRMT + IMAD alternating
			if (i & 1) asm volatile("mad.lo.u32 %0, %0, %2, %1;" : "+r"(r[i]) : "r"(s), "r"(t));
			else asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(t), "r"(s));
		}
		if (OP == 11) {                                                                                  // PRMT + LOP3 alternating (both ALU)
			if (i & 1) asm volatile("lop3.b32 %0, %0, 0x03030303, %1, 0x78;" : "+r"(r[i]) : "r"(t));
			else asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(t), "r"(s));
		}
		if (OP == 12) {                                                                                  // 2 PRMT : 1 IMAD : 1 IMAD.HI
			if ((i & 3) == 1) asm volatile("mad.lo.u32 %0, %0, %2, %1;" : "+r"(r[i]) : "r"(s), "r"(t));
			else if ((i & 3) == 3) asm volatile("mul.hi.u32 %0, %0, 65536;" : "+r"(r[i]));
			else asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(t), "r"(s));
		}
		if (OP == 13) asm volatile("mul.lo.u32 %0, %0, 0x010D0000;" : "+r"(r[i]));                       // IMAD (mul.lo imm)
		if (OP == 15) { uint32_t lo, hi; asm volatile("{ .reg .b64 w; mul.wide.u32 w, %2, %3; mov.b64 {%0, %1}, w; }" : "=r"(lo), "=r"(hi) : "r"(r[i]), "r"(t)); r[i] = hi ^ lo; }   // IMAD.WIDE by a register (+ LOP)
		if (OP == 19) { uint32_t hi; asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(hi) : "r"(r[i]), "r"(t)); r[i] = hi; }   // IMAD.HI by a register
		if (OP == 16) asm volatile("shf.r.clamp.b32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(s), "r"(t));       // SHF funnel
		if (OP == 17) asm volatile("bfe.u32 %0, %0, 16, 16;" : "+r"(r[i]));                                 // BFE
		if (OP == 18) asm volatile("shr.u32 %0, %0, %1;" : "+r"(r[i]) : "r"(t));                            // SHF by register
		if (OP == 20) asm volatile("prmt.b32 %0, %1, %2, %3;" : "=r"(r[i]) : "r"(r[(i + 1) & 7]), "r"(r[(i + 3) & 7]), "r"(r[(i + 5) & 7]));   // PRMT, 3 fresh regs
		if (OP == 21) asm volatile("prmt.b32 %0, %1, %2, %3;" : "=r"(r[i]) : "r"(r[(i + 1) & 7]), "r"(r[(i + 2) & 7]), "r"(s));                   // PRMT, 2 fresh regs
		if (OP == 22) asm volatile("lop3.b32 %0, %1, %2, %3, 0x78;" : "=r"(r[i]) : "r"(r[(i + 1) & 7]), "r"(r[(i + 3) & 7]), "r"(r[(i + 5) & 7])); // LOP3, 3 fresh regs
		if (OP == 23) asm volatile("lop3.b32 %0, %1, 0x10101010, %2, 0x78;" : "=r"(r[i]) : "r"(r[(i + 1) & 7]), "r"(r[(i + 3) & 7]));             // LOP3, 2 fresh + imm
		if (OP == 14) asm volatile("vmin4.u32.u32.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(s), "r"(t)); // emulated SIMD min
	}
}

template <int OP>
__global__ void k(uint32_t* out, long long* cycles, uint32_t s, uint32_t t) {
	uint32_t r[8];
#pragma unroll
	for (int i = 0; i < 8; ++i) r[i] = threadIdx.x * 8 + i + s;
	__syncthreads();
	const long long t0 = clock64();
	for (int it = 0; it < ITER; ++it) op8<OP>(r, s, t);
	const long long t1 = clock64();
	uint32_t acc = 0;
#pragma unroll
	for (int i = 0; i < 8; ++i) acc ^= r[i];
	out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
	if ((threadIdx.x & 31) == 0) cycles[blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32] = t1 - t0;
}

template <int OP>
void run(const char* name) {
	uint32_t* out; long long* cyc;
	cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 32 * 8);
	for (int warps_per_smsp : {1, 2, 4, 8}) {
		const int threads = warps_per_smsp * 4 * 32;
		k<OP><<<148, threads>>>(out, cyc, 0x3210, 65536 + 7);
		k<OP><<<148, threads>>>(out, cyc, 0x3210, 65536 + 7);
		cudaDeviceSynchronize();
		long long h[148 * 32];
		cudaMemcpy(h, cyc, sizeof(long long) * 148 * threads / 32, cudaMemcpyDeviceToHost);
		long long mx = 0;
		for (int i = 0; i < 148 * threads / 32; ++i) mx = h[i] > mx ? h[i] : mx;
		printf("%-28s warps/SMSP %d  inst/clk/SMSP %.3f\n", name, warps_per_smsp, (double)warps_per_smsp * ITER * 8 / (double)mx);
	}
	cudaFree(out); cudaFree(cyc);
}

int main() {
	run<0>("PRMT reg-sel");
	run<1>("PRMT imm-sel");
	run<2>("LOP3 3-reg");
	run<3>("LOP3 imm");
	run<4>("IADD");
	run<5>("IMAD imm");
	run<6>("IMAD.HI");
	run<13>("IMAD mul.lo imm");
	run<7>("IDP.4A");
	run<8>("SHF");
	run<14>("vmin4 (emulated)");
	run<20>("PRMT 3 fresh regs");
	run<21>("PRMT 2 fresh regs");
	run<22>("LOP3 3 fresh regs");
	run<23>("LOP3 2 fresh + imm");
	run<15>("IMAD.WIDE reg + LOP");
	run<19>("IMAD.HI reg");
	run<16>("SHF funnel");
	run<17>("BFE 16,16");
	run<18>("SHR by reg");
	run<9>("PRMT+IMAD.HI 1:1");
	run<10>("PRMT+IMAD 1:1");
	run<11>("PRMT+LOP3 1:1");
	run<12>("2PRMT:IMAD:IMAD.HI");
	return 0;
}
