// Shared-memory load micro-benchmark (sm_100a): cycles per warp-instruction of LDS.128 / LDS.64 / LDS.32 for the per-lane address
// patterns the packed advection kernels produce. Cells are float4 (16 B); a warp is 4 rows (quarters) of 8 z-lanes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/ubench/lds128_patterns scripts/ubench/lds128_patterns.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

constexpr int kCellsTotal = 14 * 14 * 16;
constexpr int kIters = 2048;

template <int W>  // W = words per load (1, 2, 4)
__global__ void k(const int* __restrict__ cell_of_lane, int word, long long* cycles, float* sink, uint32_t step) {
	extern __shared__ __align__(128) float4 reg[];
	for (int i = threadIdx.x; i < kCellsTotal; i += blockDim.x) reg[i] = make_float4(float(i), 1.f, 2.f, 3.f);
	__syncthreads();
	const int lane = threadIdx.x & 31;
	const int c = cell_of_lane[lane];
	uint32_t base = uint32_t(__cvta_generic_to_shared(reg)) + uint32_t(c) * 16u + uint32_t(word) * 4u;
	float acc = 0.f;
	__syncthreads();
	const long long t0 = clock64();
#pragma unroll 1
	for (int it = 0; it < kIters; ++it) {
		base += step;  // 0 at run time: keeps the loads inside the loop
#pragma unroll
		for (int u = 0; u < 8; ++u) {
			// eight independent loads per iteration; the address does not depend on the data, so the pipe stays full
			if (W == 4) {
				float x, y, z, w;
				asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x), "=f"(y), "=f"(z), "=f"(w) : "r"(base + u * step) : "memory");
				acc += (x + y) + (z + w);
			} else if (W == 2) {
				float x, y;
				asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(x), "=f"(y) : "r"(base + u * step) : "memory");
				acc += x + y;
			} else {
				float x;
				asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(base + u * step) : "memory");
				acc += x;
			}
		}
	}
	const long long t1 = clock64();
	if (threadIdx.x == 0) *cycles = t1 - t0;
	sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main() {
	struct Pat {
		const char* name;
		int cell[32];
	};
	auto lane_cell = [](int y, int z) { return (3 + y) * 16 + 4 + z + 3 * 224; };
	Pat pats[16];
	int np = 0;
	auto add = [&](const char* name, auto f) {
		pats[np].name = name;
		for (int l = 0; l < 32; ++l) pats[np].cell[l] = f(l >> 3, l & 7);
		++np;
	};
	add("aligned: lane (y,z) -> own cell", [&](int y, int z) { return lane_cell(y, z); });
	add("z pairs share a cell (duplicates inside a quarter)", [&](int y, int z) { return lane_cell(y, z & ~1); });
	add("one duplicate per quarter (lanes 3,4 share)", [&](int y, int z) { return lane_cell(y, z < 4 ? z : z - 1); });
	add("z step: lanes z>=4 read z+1 (no duplicate, no wrap)", [&](int y, int z) { return lane_cell(y, z < 4 ? z : z + 1) - 1; });
	add("wrap: lane 0 reads z-1.. lane 7 reads z+... (cells -1 and 7)", [&](int y, int z) { return lane_cell(y, z == 0 ? -1 : z); });
	add("y disagreement: odd z lanes read the next row", [&](int y, int z) { return lane_cell(y + (z & 1), z); });
	add("x disagreement: odd z lanes read the next plane", [&](int y, int z) { return lane_cell(y, z) + (z & 1) * 224; });
	add("all 32 lanes one cell", [&](int, int) { return lane_cell(0, 0); });
	add("each quarter one cell", [&](int y, int) { return lane_cell(y, 0); });
	add("random floor flips in z (cell z-1+d)", [&](int y, int z) { return lane_cell(y, z - 1 + ((0x5A3C96E1u >> (y * 8 + z)) & 1)); });
	add("random flips in x, y and z", [&](int y, int z) {
		const unsigned h = 0x9E3779B9u * unsigned(y * 8 + z + 1);
		return lane_cell(y - 1 + ((h >> 7) & 1), z - 1 + ((h >> 13) & 1)) - 224 * ((h >> 19) & 1);
	});
	add("duplicates across quarters only (rows share a cell)", [&](int y, int z) { return lane_cell(y & ~1, z); });
	int* d_cells;
	long long* d_cycles;
	float* d_sink;
	cudaMalloc(&d_cells, 32 * sizeof(int));
	cudaMalloc(&d_cycles, sizeof(long long));
	cudaMalloc(&d_sink, 1024 * sizeof(float));
	const size_t smem = kCellsTotal * sizeof(float4);
	cudaFuncSetAttribute(k<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
	cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
	cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
	printf("cycles per warp-instruction with 16 warps on one SM issuing (= shared-memory wavefronts per instruction when the pipe is the limit)\n");
	printf("%-62s %8s %8s %8s %8s\n", "pattern", "LDS.128", "LDS.64", "LDS.32", "LDS.32+12");
	for (int p = 0; p < np; ++p) {
		cudaMemcpy(d_cells, pats[p].cell, sizeof(pats[p].cell), cudaMemcpyHostToDevice);
		double r[4];
		for (int v = 0; v < 4; ++v) {
			long long c = 0;
			for (int rep = 0; rep < 2; ++rep) {
				if (v == 0) k<4><<<1, 512, smem>>>(d_cells, 0, d_cycles, d_sink, 0u);
				if (v == 1) k<2><<<1, 512, smem>>>(d_cells, 0, d_cycles, d_sink, 0u);
				if (v == 2) k<1><<<1, 512, smem>>>(d_cells, 0, d_cycles, d_sink, 0u);
				if (v == 3) k<1><<<1, 512, smem>>>(d_cells, 3, d_cycles, d_sink, 0u);
				cudaDeviceSynchronize();
			}
			cudaMemcpy(&c, d_cycles, sizeof(c), cudaMemcpyDeviceToHost);
			r[v] = double(c) / (double(kIters) * 8 * 16);  // 16 warps share the SM's one shared-memory pipe
		}
		printf("%-62s %8.2f %8.2f %8.2f %8.2f\n", pats[p].name, r[0], r[1], r[2], r[3]);
	}
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(e));
	return 0;
}
