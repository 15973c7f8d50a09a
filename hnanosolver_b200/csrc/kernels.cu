// Hand-written sm_100a kernels of the HNanoSolver hot path.
//
// Design: the reference runs one thread per voxel and resolves every neighbour with a three-level
// NanoVDB tree walk (externals/nanovdb/NanoVDB.h:5683-5698). Here the work unit is the 8^3 leaf brick: fields are
// brick-major float[L][512] (velocity as three planes), neighbour bricks come from a per-leaf 27-entry table, rows
// of eight z-consecutive voxels move as two 128-bit accesses, and the pressure sweep stages brick + halo in shared memory
// so that a red and a black half-sweep cost one pass over HBM. Arithmetic follows the reference kernels operation by
// operation, with the multiply-adds ptxas fuses in the reference build written as explicit fmaf (see oracle/hns_oracle.c).
#include "kernels.cuh"

namespace hns {

std::atomic<uint64_t> g_launches{0};

// =============================================================================================================
// small helpers
// =============================================================================================================
struct Row8 {
	float v[8];
};
__device__ __forceinline__ Row8 ld_row(const float* __restrict__ f, uint64_t idx) {
	const float4 a = __ldg(reinterpret_cast<const float4*>(f + idx));
	const float4 b = __ldg(reinterpret_cast<const float4*>(f + idx) + 1);
	return Row8{{a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w}};
}
// plain (coherent) loads for fields that other CTAs of the same launch write (in-place half-sweeps)
__device__ __forceinline__ Row8 ld_row_coherent(const float* f, uint64_t idx) {
	const float4 a = *reinterpret_cast<const float4*>(f + idx);
	const float4 b = *(reinterpret_cast<const float4*>(f + idx) + 1);
	return Row8{{a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w}};
}
__device__ __forceinline__ void st_row(float* __restrict__ f, uint64_t idx, const Row8& r) {
	reinterpret_cast<float4*>(f + idx)[0] = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
	reinterpret_cast<float4*>(f + idx)[1] = make_float4(r.v[4], r.v[5], r.v[6], r.v[7]);
}
__device__ __forceinline__ Row8 zero_row() { return Row8{{0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}}; }

// Row-per-thread addressing: 64 threads per leaf, thread r owns voxels (x = r>>3, y = r&7, z = 0..7).
struct RowCtx {
	uint32_t leaf;
	int x, y;
	const int32_t* nbr;  // this leaf's 27-entry table
	__device__ __forceinline__ uint64_t self() const { return uint64_t(leaf) * 512u + uint32_t(x * 64 + y * 8); }
	// index of row (x+dx, y+dy) with dx,dy in {-1,0,1} (one of them 0), or -1 when it lies in a missing leaf
	__device__ __forceinline__ int64_t row(int dx, int dy) const {
		int xx = x + dx, yy = y + dy;
		int slot = kSlotSelf;
		if (xx < 0) slot = kSlotXm, xx = 7;
		else if (xx > 7) slot = kSlotXp, xx = 0;
		if (yy < 0) slot = kSlotYm, yy = 7;
		else if (yy > 7) slot = kSlotYp, yy = 0;
		const int32_t l = slot == kSlotSelf ? int32_t(leaf) : __ldg(nbr + slot);
		return l < 0 ? int64_t(-1) : int64_t(uint64_t(l) * 512u + uint32_t(xx * 64 + yy * 8));
	}
	// index of voxel (x, y, z = 7) of the -z neighbour leaf / (x, y, z = 0) of the +z neighbour leaf, or -1
	__device__ __forceinline__ int64_t zminus() const {
		const int32_t l = __ldg(nbr + kSlotZm);
		return l < 0 ? int64_t(-1) : int64_t(uint64_t(l) * 512u + uint32_t(x * 64 + y * 8 + 7));
	}
	__device__ __forceinline__ int64_t zplus() const {
		const int32_t l = __ldg(nbr + kSlotZp);
		return l < 0 ? int64_t(-1) : int64_t(uint64_t(l) * 512u + uint32_t(x * 64 + y * 8));
	}
};
__device__ __forceinline__ bool make_row_ctx(const GridView& g, RowCtx& c) {
	const uint32_t leaf = blockIdx.x * (blockDim.x >> 6) + (threadIdx.x >> 6);
	if (leaf >= g.num_leaves) return false;
	const int r = threadIdx.x & 63;
	c.leaf = leaf, c.x = r >> 3, c.y = r & 7, c.nbr = g.nbr + uint64_t(leaf) * 27u;
	return true;
}

// =============================================================================================================
// layout conversion
// =============================================================================================================
__global__ void __launch_bounds__(256) k_aos_to_soa(const float* __restrict__ aos, float* __restrict__ u, float* __restrict__ v,
                                                    float* __restrict__ w, uint64_t n) {
	const uint64_t t = blockIdx.x * uint64_t(256) + threadIdx.x;
	if (t >= n) return;
	u[t] = __ldg(aos + 3 * t), v[t] = __ldg(aos + 3 * t + 1), w[t] = __ldg(aos + 3 * t + 2);
}
__global__ void __launch_bounds__(256) k_soa_to_aos(const float* __restrict__ u, const float* __restrict__ v, const float* __restrict__ w,
                                                    float* __restrict__ aos, uint64_t n) {
	const uint64_t t = blockIdx.x * uint64_t(256) + threadIdx.x;
	if (t >= n) return;
	aos[3 * t] = __ldg(u + t), aos[3 * t + 1] = __ldg(v + t), aos[3 * t + 2] = __ldg(w + t);
}
void launch_aos_to_soa(const float* aos, float* u, float* v, float* w, uint64_t n, cudaStream_t st) {
	if (n) HNS_LAUNCH(k_aos_to_soa, unsigned((n + 255) / 256), 256, 0, st, aos, u, v, w, n);
}
void launch_soa_to_aos(const float* u, const float* v, const float* w, float* aos, uint64_t n, cudaStream_t st) {
	if (n) HNS_LAUNCH(k_soa_to_aos, unsigned((n + 255) / 256), 256, 0, st, u, v, w, aos, n);
}

// =============================================================================================================
// divergence  (reference Kernel.cu:499-519)
//   xp = (c.x + u(+x).x) * 0.5 ... ; div = (xp - xm + yp - ym + zp - zm) * inv_dx ; inactive neighbour -> 0
// =============================================================================================================
__global__ void __launch_bounds__(256) k_divergence(GridView g, const float* __restrict__ u, const float* __restrict__ v,
                                                    const float* __restrict__ w, float* __restrict__ div, float inv_dx) {
	RowCtx c;
	if (!make_row_ctx(g, c)) return;
	const uint64_t self = c.self();
	const Row8 cu = ld_row(u, self), cv = ld_row(v, self), cw = ld_row(w, self);
	int64_t i;
	const Row8 uxp = (i = c.row(1, 0)) >= 0 ? ld_row(u, i) : zero_row();
	const Row8 uxm = (i = c.row(-1, 0)) >= 0 ? ld_row(u, i) : zero_row();
	const Row8 vyp = (i = c.row(0, 1)) >= 0 ? ld_row(v, i) : zero_row();
	const Row8 vym = (i = c.row(0, -1)) >= 0 ? ld_row(v, i) : zero_row();
	const float wzm = (i = c.zminus()) >= 0 ? __ldg(w + i) : 0.f;
	const float wzp = (i = c.zplus()) >= 0 ? __ldg(w + i) : 0.f;
	Row8 o;
#pragma unroll
	for (int z = 0; z < 8; ++z) {
		const float xp = (cu.v[z] + uxp.v[z]) * 0.5f, xm = (cu.v[z] + uxm.v[z]) * 0.5f;
		const float yp = (cv.v[z] + vyp.v[z]) * 0.5f, ym = (cv.v[z] + vym.v[z]) * 0.5f;
		const float zp = (cw.v[z] + (z < 7 ? cw.v[z < 7 ? z + 1 : 7] : wzp)) * 0.5f;
		const float zm = (cw.v[z] + (z > 0 ? cw.v[z > 0 ? z - 1 : 0] : wzm)) * 0.5f;
		o.v[z] = (xp - xm + yp - ym + zp - zm) * inv_dx;
	}
	st_row(div, self, o);
}
void launch_divergence(const GridView& g, const float* const vel[3], float* div, float inv_dx, cudaStream_t st) {
	if (g.num_leaves) HNS_LAUNCH(k_divergence, (g.num_leaves + 3) / 4, 256, 0, st, g, vel[0], vel[1], vel[2], div, inv_dx);
}

// =============================================================================================================
// subtractPressureGradient  (reference Kernel.cu:765-829): u - ((p(+1) - p(-1)) * 0.5) * inv_dx, fused as in the reference SASS
// =============================================================================================================
__global__ void __launch_bounds__(256) k_subtract_gradient(GridView g, const float* __restrict__ u, const float* __restrict__ v,
                                                           const float* __restrict__ w, const float* __restrict__ p, float* __restrict__ ou,
                                                           float* __restrict__ ov, float* __restrict__ ow, float inv_dx) {
	RowCtx c;
	if (!make_row_ctx(g, c)) return;
	const uint64_t self = c.self();
	const Row8 cp = ld_row(p, self);
	int64_t i;
	const Row8 pxp = (i = c.row(1, 0)) >= 0 ? ld_row(p, i) : zero_row();
	const Row8 pxm = (i = c.row(-1, 0)) >= 0 ? ld_row(p, i) : zero_row();
	const Row8 pyp = (i = c.row(0, 1)) >= 0 ? ld_row(p, i) : zero_row();
	const Row8 pym = (i = c.row(0, -1)) >= 0 ? ld_row(p, i) : zero_row();
	const float pzm = (i = c.zminus()) >= 0 ? __ldg(p + i) : 0.f;
	const float pzp = (i = c.zplus()) >= 0 ? __ldg(p + i) : 0.f;
	const Row8 cu = ld_row(u, self), cv = ld_row(v, self), cw = ld_row(w, self);
	Row8 a, b, d;
#pragma unroll
	for (int z = 0; z < 8; ++z) {
		const float zp = z < 7 ? cp.v[z < 7 ? z + 1 : 7] : pzp;
		const float zm = z > 0 ? cp.v[z > 0 ? z - 1 : 0] : pzm;
		a.v[z] = fmaf(-((pxp.v[z] - pxm.v[z]) * 0.5f), inv_dx, cu.v[z]);
		b.v[z] = fmaf(-((pyp.v[z] - pym.v[z]) * 0.5f), inv_dx, cv.v[z]);
		d.v[z] = fmaf(-((zp - zm) * 0.5f), inv_dx, cw.v[z]);
	}
	st_row(ou, self, a);
	st_row(ov, self, b);
	st_row(ow, self, d);
}
void launch_subtract_gradient(const GridView& g, const float* const vel[3], const float* p, float* const out[3], float inv_dx, cudaStream_t st) {
	if (g.num_leaves)
		HNS_LAUNCH(k_subtract_gradient, (g.num_leaves + 3) / 4, 256, 0, st, g, vel[0], vel[1], vel[2], p, out[0], out[1], out[2], inv_dx);
}

// =============================================================================================================
// red-black Gauss-Seidel / SOR  (reference Kernel.cu:591-623)
//   s = (pxp + pxm + pyp + pym + pzp + pzm) - div*dx^2 ; pGS = s/6 ; p = pOld + omega*(pGS - pOld)
//   as compiled in the reference: s = fma(-div, dx2, sum); d = fma(s, 1/6, -pOld); p = fma(d, omega, pOld)
// =============================================================================================================
__device__ __forceinline__ float sor_update(float pxp, float pxm, float pyp, float pym, float pzp, float pzm, float dv, float pOld, float dx2,
                                            float omega) {
	const float s = fmaf(-dv, dx2, ((((pxp + pxm) + pyp) + pym) + pzp) + pzm);
	const float d = fmaf(s, 0.166666667f, -pOld);
	return fmaf(d, omega, pOld);
}

// --- one colour per launch, in place (the reference's own schedule; kept as fallback and as cross-check of the fused kernel) ---
__global__ void __launch_bounds__(256) k_rbgs_color(GridView g, const float* __restrict__ div, float* p, float dx2, int color, float omega) {
	RowCtx c;
	if (!make_row_ctx(g, c)) return;
	const uint64_t self = c.self();
	Row8 cp = ld_row_coherent(p, self);
	int64_t i;
	const Row8 pxp = (i = c.row(1, 0)) >= 0 ? ld_row_coherent(p, i) : zero_row();
	const Row8 pxm = (i = c.row(-1, 0)) >= 0 ? ld_row_coherent(p, i) : zero_row();
	const Row8 pyp = (i = c.row(0, 1)) >= 0 ? ld_row_coherent(p, i) : zero_row();
	const Row8 pym = (i = c.row(0, -1)) >= 0 ? ld_row_coherent(p, i) : zero_row();
	const float pzm = (i = c.zminus()) >= 0 ? p[i] : 0.f;
	const float pzp = (i = c.zplus()) >= 0 ? p[i] : 0.f;
	const Row8 dv = ld_row(div, self);
	const int s0 = (c.x + c.y + color) & 1;  // parity of the z's of this colour in the row (leaf origins are multiples of 8)
	Row8 o;
#pragma unroll
	for (int z = 0; z < 8; ++z) {
		const float zp = z < 7 ? cp.v[z < 7 ? z + 1 : 7] : pzp;
		const float zm = z > 0 ? cp.v[z > 0 ? z - 1 : 0] : pzm;
		const float r = sor_update(pxp.v[z], pxm.v[z], pyp.v[z], pym.v[z], zp, zm, dv.v[z], cp.v[z], dx2, omega);
		o.v[z] = (z & 1) == s0 ? r : cp.v[z];
	}
	// the other colour is stored back unchanged: nobody updates it during this launch, so concurrent readers see the same value
	st_row(p, self, o);
}
void launch_rbgs_color(const GridView& g, const float* div, float* p, float dx, int color, float omega, cudaStream_t st) {
	if (g.num_leaves) HNS_LAUNCH(k_rbgs_color, (g.num_leaves + 3) / 4, 256, 0, st, g, div, p, dx * dx, color, omega);
}

// --- red + black in one launch --------------------------------------------------------------------------------
// One CTA (256 threads) per leaf. The brick's pressure plus a two-voxel halo is staged in a 12^3 shared-memory cube:
//   sweep 0 updates the red voxels of the brick AND the red voxels of the one-voxel face ring (recomputing what the
//   neighbouring CTAs compute for their own bricks), which is exactly what the black voxels of the brick need for
//   sweep 1. Every update evaluates the same expression on the same operands as the two-launch schedule, so results
//   are bit-identical; HBM traffic per full iteration drops from 2 x (read p, read div, write p) to 1 x.
//   Reads p_in, writes p_out (other CTAs still need the old halo values, so the sweep cannot be in place).
constexpr int kCube = 12;
constexpr int kCubeN = kCube * kCube * kCube;
constexpr int kHaloCells = 6 * 128 + 12 * 8;  // six 2-deep face slabs + twelve 1-voxel edges
constexpr int kRingCells = 6 * 32;            // red voxels of the six one-voxel face slabs
// halo entry: bits 0-10 cube index, 11-15 neighbour slot, 16-24 source voxel offset
__device__ uint32_t d_halo_tab[kHaloCells];
__device__ uint32_t d_ring_tab[kRingCells];
__host__ __device__ constexpr int cube_idx(int x, int y, int z) { return ((x + 2) * kCube + (y + 2)) * kCube + (z + 2); }

int upload_tables() {
	static uint32_t halo[kHaloCells], ring[kRingCells];
	int nh = 0, nr = 0;
	auto entry = [](int x, int y, int z) {  // region coordinate (-2..9) -> packed entry
		const int dx = x < 0 ? -1 : (x > 7 ? 1 : 0), dy = y < 0 ? -1 : (y > 7 ? 1 : 0), dz = z < 0 ? -1 : (z > 7 ? 1 : 0);
		const uint32_t slot = uint32_t((dx + 1) * 9 + (dy + 1) * 3 + (dz + 1));
		const uint32_t src = uint32_t(((x & 7) << 6) | ((y & 7) << 3) | (z & 7));
		return uint32_t(cube_idx(x, y, z)) | slot << 11 | src << 16;
	};
	for (int axis = 0; axis < 3; ++axis)
		for (int side = 0; side < 2; ++side)
			for (int depth = 1; depth <= 2; ++depth)
				for (int a = 0; a < 8; ++a)
					for (int b = 0; b < 8; ++b) {
						const int h = side ? 7 + depth : -depth;
						const int x = axis == 0 ? h : a, y = axis == 1 ? h : (axis == 0 ? a : b), z = axis == 2 ? h : b;
						halo[nh++] = entry(x, y, z);
						if (depth == 1 && ((x + y + z) & 1) == 0) ring[nr++] = entry(x, y, z);
					}
	for (int axis = 0; axis < 3; ++axis)  // edges parallel to `axis`
		for (int s1 = 0; s1 < 2; ++s1)
			for (int s2 = 0; s2 < 2; ++s2)
				for (int a = 0; a < 8; ++a) {
					const int h1 = s1 ? 8 : -1, h2 = s2 ? 8 : -1;
					const int x = axis == 0 ? a : h1, y = axis == 1 ? a : (axis == 0 ? h1 : h2), z = axis == 2 ? a : h2;
					halo[nh++] = entry(x, y, z);
				}
	if (nh != kHaloCells || nr != kRingCells) return fail(HNS_ERR_RUNTIME, "internal: halo table size");
	HNS_CUDA(cudaMemcpyToSymbol(d_halo_tab, halo, sizeof(halo)));
	HNS_CUDA(cudaMemcpyToSymbol(d_ring_tab, ring, sizeof(ring)));
	return HNS_OK;
}

__global__ void __launch_bounds__(256) k_rbgs_fused(GridView g, const float* __restrict__ div, const float* __restrict__ p_in,
                                                    float* __restrict__ p_out, float dx2, float omega) {
	__shared__ __align__(16) float cube[kCubeN];
	__shared__ int32_t s_nbr[27];
	const uint32_t leaf = blockIdx.x;
	const int tid = threadIdx.x;
	if (tid < 27) s_nbr[tid] = __ldg(g.nbr + uint64_t(leaf) * 27u + tid);
	// own pair: row r = tid>>2 -> (x, y), z = 2j, 2j+1
	const int x = tid >> 5, y = (tid >> 2) & 7, j = tid & 3;
	const uint64_t self = uint64_t(leaf) * 512u + uint32_t(tid * 2);
	const float2 pp = __ldg(reinterpret_cast<const float2*>(p_in + self));
	const float2 dd = __ldg(reinterpret_cast<const float2*>(div + self));
	const int ci = cube_idx(x, y, 2 * j);
	*reinterpret_cast<float2*>(&cube[ci]) = pp;
	__syncthreads();  // s_nbr visible
	// halo
	for (int h = tid; h < kHaloCells; h += 256) {
		const uint32_t e = d_halo_tab[h];
		const int32_t l = s_nbr[(e >> 11) & 31u];
		cube[e & 2047u] = l < 0 ? 0.f : __ldg(p_in + uint64_t(l) * 512u + (e >> 16));
	}
	// ring voxel handled by this thread in sweep 0 (threads 0..191)
	int ring_ci = -1;
	float ring_div = 0.f;
	if (tid < kRingCells) {
		const uint32_t e = d_ring_tab[tid];
		const int32_t l = s_nbr[(e >> 11) & 31u];
		if (l >= 0) {
			ring_ci = int(e & 2047u);
			ring_div = __ldg(div + uint64_t(l) * 512u + (e >> 16));
		}
	}
	__syncthreads();
	// ---- sweep 0: red = (x+y+z) even ----
	const int s = (x + y) & 1;          // 0: z = 2j is red, 1: z = 2j+1 is red
	const int cr = ci + s, cb = ci + (s ^ 1);
	const float pr_old = s ? pp.y : pp.x, pb_old = s ? pp.x : pp.y;
	const float dr = s ? dd.y : dd.x, db = s ? dd.x : dd.y;
	const float pr = sor_update(cube[cr + 144], cube[cr - 144], cube[cr + 12], cube[cr - 12], cube[cr + 1], cube[cr - 1], dr, pr_old, dx2, omega);
	float ring_new = 0.f;
	if (ring_ci >= 0) {
		const int c = ring_ci;
		ring_new = sor_update(cube[c + 144], cube[c - 144], cube[c + 12], cube[c - 12], cube[c + 1], cube[c - 1], ring_div, cube[c], dx2, omega);
	}
	// red updates read black cells only (plus their own old value), so they can be written back without a barrier
	cube[cr] = pr;
	if (ring_ci >= 0) cube[ring_ci] = ring_new;
	__syncthreads();
	// ---- sweep 1: black ----
	const float pb = sor_update(cube[cb + 144], cube[cb - 144], cube[cb + 12], cube[cb - 12], cube[cb + 1], cube[cb - 1], db, pb_old, dx2, omega);
	*reinterpret_cast<float2*>(p_out + self) = s ? make_float2(pb, pr) : make_float2(pr, pb);
}
void launch_rbgs_fused(const GridView& g, const float* div, const float* p_in, float* p_out, float dx, float omega, cudaStream_t st) {
	if (g.num_leaves) HNS_LAUNCH(k_rbgs_fused, g.num_leaves, 256, 0, st, g, div, p_in, p_out, dx * dx, omega);
}

// =============================================================================================================
// semi-Lagrangian BFECC advection  (reference Kernel.cu:118-453, samplers src/Utils/Stencils.hpp:25-173)
// =============================================================================================================
// Resolves voxel (i,j,k) (global coordinates) to a sidecar index, or -1 when inactive. The 3x3x3 leaf neighbourhood of the
// CTA's leaf comes from shared memory; anything farther away walks the NanoVDB buffer.
struct LeafFrame {
	int ox, oy, oz;
	const int32_t* s_nbr;  // shared
};
__device__ __forceinline__ int64_t voxel_index(const GridView& g, const LeafFrame& f, int i, int j, int k) {
	const int rx = i - f.ox, ry = j - f.oy, rz = k - f.oz;
	const int dx = rx >> 3, dy = ry >> 3, dz = rz >> 3;
	int32_t l;
	if (((dx + 1) | (dy + 1) | (dz + 1)) & ~3 || dx == 2 || dy == 2 || dz == 2) {  // outside the 3x3x3 neighbourhood
		l = probe_leaf(g, i, j, k);
	} else {
		l = f.s_nbr[(dx + 1) * 9 + (dy + 1) * 3 + (dz + 1)];
	}
	return l < 0 ? int64_t(-1) : int64_t(uint64_t(l) * 512u + uint32_t(((rx & 7) << 6) | ((ry & 7) << 3) | (rz & 7)));
}
__device__ __forceinline__ float lerpf(float a, float b, float w) { return fmaf(w, b - a, a); }

// TrilinearSampler<Vec3f>::sample: Floor (round down, fractional part in place), 8 nearest fetches (inactive -> 0),
// lerp z, then y, then x (Stencils.hpp:96-157)
__device__ __forceinline__ void trilinear_vec(const GridView& g, const LeafFrame& f, const float* __restrict__ u, const float* __restrict__ v,
                                              const float* __restrict__ w, float px, float py, float pz, float& ru, float& rv, float& rw) {
	const int i = __float2int_rd(px), j = __float2int_rd(py), k = __float2int_rd(pz);
	const float fx = px - float(i), fy = py - float(j), fz = pz - float(k);
	float cu[8], cv[8], cw[8];
#pragma unroll
	for (int q = 0; q < 8; ++q) {  // q = a*4 + b*2 + c  <->  v[a][b][c]
		const int64_t idx = voxel_index(g, f, i + (q >> 2), j + ((q >> 1) & 1), k + (q & 1));
		cu[q] = idx < 0 ? 0.f : __ldg(u + idx);
		cv[q] = idx < 0 ? 0.f : __ldg(v + idx);
		cw[q] = idx < 0 ? 0.f : __ldg(w + idx);
	}
	ru = lerpf(lerpf(lerpf(cu[0], cu[1], fz), lerpf(cu[2], cu[3], fz), fy), lerpf(lerpf(cu[4], cu[5], fz), lerpf(cu[6], cu[7], fz), fy), fx);
	rv = lerpf(lerpf(lerpf(cv[0], cv[1], fz), lerpf(cv[2], cv[3], fz), fy), lerpf(lerpf(cv[4], cv[5], fz), lerpf(cv[6], cv[7], fz), fy), fx);
	rw = lerpf(lerpf(lerpf(cw[0], cw[1], fz), lerpf(cw[2], cw[3], fz), fy), lerpf(lerpf(cw[4], cw[5], fz), lerpf(cw[6], cw[7], fz), fy), fx);
}
__device__ __forceinline__ float trilinear_f(const GridView& g, const LeafFrame& f, const float* __restrict__ a, float px, float py, float pz) {
	const int i = __float2int_rd(px), j = __float2int_rd(py), k = __float2int_rd(pz);
	const float fx = px - float(i), fy = py - float(j), fz = pz - float(k);
	float c[8];
#pragma unroll
	for (int q = 0; q < 8; ++q) {
		const int64_t idx = voxel_index(g, f, i + (q >> 2), j + ((q >> 1) & 1), k + (q & 1));
		c[q] = idx < 0 ? 0.f : __ldg(a + idx);
	}
	return lerpf(lerpf(lerpf(c[0], c[1], fz), lerpf(c[2], c[3], fz), fy), lerpf(lerpf(c[4], c[5], fz), lerpf(c[6], c[7], fz), fy), fx);
}

// One CTA of 512 threads per leaf, one thread per voxel.
__device__ __forceinline__ void leaf_frame(const GridView& g, int32_t* s_nbr, LeafFrame& f, int& x, int& y, int& z) {
	const uint32_t leaf = blockIdx.x;
	if (threadIdx.x < 27) s_nbr[threadIdx.x] = __ldg(g.nbr + uint64_t(leaf) * 27u + threadIdx.x);
	const int4 o = __ldg(g.origin + leaf);
	f.ox = o.x, f.oy = o.y, f.oz = o.z, f.s_nbr = s_nbr;
	x = threadIdx.x >> 6, y = (threadIdx.x >> 3) & 7, z = threadIdx.x & 7;
	__syncthreads();
}

__global__ void __launch_bounds__(512) k_advect_vector(GridView g, const float* __restrict__ u, const float* __restrict__ v,
                                                       const float* __restrict__ w, float* __restrict__ ou, float* __restrict__ ov,
                                                       float* __restrict__ ow, float sdt) {
	__shared__ int32_t s_nbr[27];
	LeafFrame f;
	int x, y, z;
	leaf_frame(g, s_nbr, f, x, y, z);
	const uint64_t self = uint64_t(blockIdx.x) * 512u + threadIdx.x;
	const int ci = f.ox + x, cj = f.oy + y, ck = f.oz + z;
	const float u0 = __ldg(u + self), v0 = __ldg(v + self), w0 = __ldg(w + self);
	// backtrace: pos - velOrig * scaled_dt  (Kernel.cu:374)
	const float bx = fmaf(-sdt, u0, float(ci)), by = fmaf(-sdt, v0, float(cj)), bz = fmaf(-sdt, w0, float(ck));
	float uf, vf, wf, ub, vb, wb;
	trilinear_vec(g, f, u, v, w, bx, by, bz, uf, vf, wf);
	const float fx = fmaf(sdt, uf, bx), fy = fmaf(sdt, vf, by), fz = fmaf(sdt, wf, bz);  // :387
	trilinear_vec(g, f, u, v, w, fx, fy, fz, ub, vb, wb);
	float cu = fmaf(0.5f, u0 - ub, uf), cv = fmaf(0.5f, v0 - vb, vf), cw = fmaf(0.5f, w0 - wb, wf);  // :399-400
	float mnu = u0, mxu = u0, mnv = v0, mxv = v0, mnw = w0, mxw = w0;
#pragma unroll
	for (int q = 0; q < 6; ++q) {  // -x, +x, -y, +y, -z, +z  (:410-421)
		const int d = (q & 1) ? 1 : -1;
		const int64_t idx = voxel_index(g, f, ci + (q < 2 ? d : 0), cj + ((q >> 1) == 1 ? d : 0), ck + (q >= 4 ? d : 0));
		const float nu = idx < 0 ? 0.f : __ldg(u + idx), nv = idx < 0 ? 0.f : __ldg(v + idx), nw = idx < 0 ? 0.f : __ldg(w + idx);
		mnu = fminf(mnu, nu), mxu = fmaxf(mxu, nu);
		mnv = fminf(mnv, nv), mxv = fmaxf(mxv, nv);
		mnw = fminf(mnw, nw), mxw = fmaxf(mxw, nw);
	}
	mnu = fminf(mnu, uf), mxu = fmaxf(mxu, uf);
	mnv = fminf(mnv, vf), mxv = fmaxf(mxv, vf);
	mnw = fminf(mnw, wf), mxw = fmaxf(mxw, wf);
	ou[self] = fmaxf(mnu, fminf(cu, mxu));  // :429
	ov[self] = fmaxf(mnv, fminf(cv, mxv));
	ow[self] = fmaxf(mnw, fminf(cw, mxw));
}
void launch_advect_vector(const GridView& g, const float* const vel[3], float* const out[3], float dt, float inv_dx, cudaStream_t st) {
	if (g.num_leaves) HNS_LAUNCH(k_advect_vector, g.num_leaves, 512, 0, st, g, vel[0], vel[1], vel[2], out[0], out[1], out[2], dt * inv_dx);
}

// advect_scalars (Kernel.cu:118-266): explicit corner weights, fma accumulation in corner order
// (i0j0k0),(i1j0k0),(i0j1k0),(i1j1k0),(i0j0k1),...; inactive corner / neighbour -> array element 0 (:192,:225).
struct Interp {
	uint32_t idx[8];
	float w[8];
};
__device__ __forceinline__ void setup_interp(const GridView& g, const LeafFrame& f, float px, float py, float pz, Interp& d) {
	const int i0 = __float2int_rd(px), j0 = __float2int_rd(py), k0 = __float2int_rd(pz);
	const float tx = px - float(i0), ty = py - float(j0), tz = pz - float(k0);
	const float itx = 1.0f - tx, ity = 1.0f - ty, itz = 1.0f - tz;
	const float w00 = itx * ity, w10 = tx * ity, w01 = itx * ty, w11 = tx * ty;
	d.w[0] = w00 * itz, d.w[1] = w10 * itz, d.w[2] = w01 * itz, d.w[3] = w11 * itz;
	d.w[4] = w00 * tz, d.w[5] = w10 * tz, d.w[6] = w01 * tz, d.w[7] = w11 * tz;
#pragma unroll
	for (int q = 0; q < 8; ++q) {  // q bit0 -> i, bit1 -> j, bit2 -> k
		const int64_t idx = voxel_index(g, f, i0 + (q & 1), j0 + ((q >> 1) & 1), k0 + (q >> 2));
		d.idx[q] = idx < 0 ? 0u : uint32_t(idx);
	}
}

template <int kSemantics>
__global__ void __launch_bounds__(512) k_advect_scalars(GridView g, const float* __restrict__ u, const float* __restrict__ v,
                                                        const float* __restrict__ w, ScalarPtrs sp, int S, float sdt) {
	__shared__ int32_t s_nbr[27];
	LeafFrame f;
	int x, y, z;
	leaf_frame(g, s_nbr, f, x, y, z);
	const uint64_t self = uint64_t(blockIdx.x) * 512u + threadIdx.x;
	const int ci = f.ox + x, cj = f.oy + y, ck = f.oz + z;
	const float u0 = __ldg(u + self), v0 = __ldg(v + self), w0 = __ldg(w + self);
	const float bx = fmaf(-sdt, u0, float(ci)), by = fmaf(-sdt, v0, float(cj)), bz = fmaf(-sdt, w0, float(ck));
	if (kSemantics == 0) {
		Interp B, F;
		setup_interp(g, f, bx, by, bz, B);
		float uf = 0.f, vf = 0.f, wf = 0.f;
#pragma unroll
		for (int q = 0; q < 8; ++q) {  // :201-206
			uf = fmaf(B.w[q], __ldg(u + B.idx[q]), uf);
			vf = fmaf(B.w[q], __ldg(v + B.idx[q]), vf);
			wf = fmaf(B.w[q], __ldg(w + B.idx[q]), wf);
		}
		setup_interp(g, f, fmaf(sdt, uf, bx), fmaf(sdt, vf, by), fmaf(sdt, wf, bz), F);  // :208,216
		uint32_t nb[6];
#pragma unroll
		for (int q = 0; q < 6; ++q) {  // {-1,0,0},{1,0,0},{0,-1,0},{0,1,0},{0,0,-1},{0,0,1}  (:219-226)
			const int d = (q & 1) ? 1 : -1;
			const int64_t idx = voxel_index(g, f, ci + (q < 2 ? d : 0), cj + ((q >> 1) == 1 ? d : 0), ck + (q >= 4 ? d : 0));
			nb[q] = idx < 0 ? 0u : uint32_t(idx);
		}
		for (int s = 0; s < S; ++s) {  // :229-265
			const float* __restrict__ a = sp.in[s];
			const float phi0 = __ldg(a + self);
			float phiF = 0.f, phiB = 0.f;
#pragma unroll
			for (int q = 0; q < 8; ++q) {
				phiF = fmaf(__ldg(a + B.idx[q]), B.w[q], phiF);
				phiB = fmaf(__ldg(a + F.idx[q]), F.w[q], phiB);
			}
			const float corr = fmaf(0.5f, phi0 - phiB, phiF);
			float mn = phi0, mx = phi0;
#pragma unroll
			for (int q = 0; q < 6; ++q) {
				const float val = __ldg(a + nb[q]);
				mn = fminf(mn, val), mx = fmaxf(mx, val);
			}
			mn = fminf(mn, phiF), mx = fmaxf(mx, phiF);
			sp.out[s][self] = fmaxf(mn, fminf(corr, mx));
		}
	} else {
		// advect_scalar (Kernel.cu:269-352): IndexSampler<float,1> everywhere, inactive -> 0, z-y-x lerps
		float uf, vf, wf;
		trilinear_vec(g, f, u, v, w, bx, by, bz, uf, vf, wf);
		const float fx = fmaf(sdt, uf, bx), fy = fmaf(sdt, vf, by), fz = fmaf(sdt, wf, bz);
		int64_t nb[6];
#pragma unroll
		for (int q = 0; q < 6; ++q) {
			const int d = (q & 1) ? 1 : -1;
			nb[q] = voxel_index(g, f, ci + (q < 2 ? d : 0), cj + ((q >> 1) == 1 ? d : 0), ck + (q >= 4 ? d : 0));
		}
		for (int s = 0; s < S; ++s) {
			const float* __restrict__ a = sp.in[s];
			const float phi0 = __ldg(a + self);
			const float phiF = trilinear_f(g, f, a, bx, by, bz);
			const float phiB = trilinear_f(g, f, a, fx, fy, fz);
			const float corr = fmaf(0.5f, phi0 - phiB, phiF);
			float mn = phi0, mx = phi0;
#pragma unroll
			for (int q = 0; q < 6; ++q) {
				const float val = nb[q] < 0 ? 0.f : __ldg(a + nb[q]);
				mn = fminf(mn, val), mx = fmaxf(mx, val);
			}
			mn = fminf(mn, phiF), mx = fmaxf(mx, phiF);
			sp.out[s][self] = fmaxf(mn, fminf(corr, mx));
		}
	}
}
void launch_advect_scalars(const GridView& g, const float* const vel[3], const ScalarPtrs& sp, int S, float dt, float inv_dx,
                           int sampler_semantics, cudaStream_t st) {
	if (!g.num_leaves || S <= 0) return;
	if (sampler_semantics == 0)
		HNS_LAUNCH(k_advect_scalars<0>, g.num_leaves, 512, 0, st, g, vel[0], vel[1], vel[2], sp, S, dt * inv_dx);
	else
		HNS_LAUNCH(k_advect_scalars<1>, g.num_leaves, 512, 0, st, g, vel[0], vel[1], vel[2], sp, S, dt * inv_dx);
}

// =============================================================================================================
// combustion_oxygen / temperature_buoyancy  (reference Kernel.cu:923-966, 831-847) -- element-wise
// =============================================================================================================
__global__ void __launch_bounds__(256) k_combustion_oxygen(const float* __restrict__ fuel, const float* __restrict__ waste,
                                                           const float* __restrict__ temp, float* __restrict__ div,
                                                           const float* __restrict__ flame, float* __restrict__ oFuel,
                                                           float* __restrict__ oWaste, float* __restrict__ oTemp, float* __restrict__ oFlame,
                                                           float temp_gain, float expansion, uint64_t n) {
	const uint64_t t = blockIdx.x * uint64_t(256) + threadIdx.x;
	if (t >= n) return;
	float f = fuel[t];
	const float wv = waste[t], T = temp[t], fl = flame[t];
	if (f < 0.001f) f = 0.0f;
	const float oxygen = 1.0f - f - wv;
	if (oxygen < 0.0f) {
		oFuel[t] = f, oWaste[t] = wv, oTemp[t] = T, oFlame[t] = fl;
		return;
	}
	const float burn = fminf(oxygen, f);
	oFuel[t] = f - burn;
	oWaste[t] = fmaf(burn, 2.0f, wv);
	oFlame[t] = fmaxf(fl, fminf(1.0f, burn * 10.0f));
	oTemp[t] = fmaf(burn, temp_gain, T);
	div[t] = fmaf(burn, expansion, div[t]);
}
void launch_combustion_oxygen(const float* fuel, const float* waste, const float* temp, float* div, const float* flame, float* oFuel,
                              float* oWaste, float* oTemp, float* oFlame, float temp_gain, float expansion, uint64_t n, cudaStream_t st) {
	if (n)
		HNS_LAUNCH(k_combustion_oxygen, unsigned((n + 255) / 256), 256, 0, st, fuel, waste, temp, div, flame, oFuel, oWaste, oTemp, oFlame,
		           temp_gain, expansion, n);
}
__global__ void __launch_bounds__(256) k_buoyancy(float* __restrict__ u, float* __restrict__ v, float* __restrict__ w,
                                                  const float* __restrict__ temp, float dt, float ambient, float strength, uint64_t n) {
	const uint64_t t = blockIdx.x * uint64_t(256) + threadIdx.x;
	if (t >= n) return;
	const float T = temp[t];
	if (T <= ambient) return;  // reference copies the velocity through unchanged (in == out there)
	const float b = fmaxf(0.0f, (T - ambient) * strength);
	u[t] = fmaf(0.0f, dt, u[t]);
	v[t] = fmaf(b, dt, v[t]);
	w[t] = fmaf(0.0f, dt, w[t]);
}
void launch_buoyancy(float* const vel[3], const float* temp, float dt, float ambient, float strength, uint64_t n, cudaStream_t st) {
	if (n) HNS_LAUNCH(k_buoyancy, unsigned((n + 255) / 256), 256, 0, st, vel[0], vel[1], vel[2], temp, dt, ambient, strength, n);
}

// =============================================================================================================
// brick gather / scatter for ghost-leaf exchange
// =============================================================================================================
__global__ void __launch_bounds__(128) k_pack_leaves(const float* __restrict__ field, const int32_t* __restrict__ ids, float* __restrict__ dst) {
	const int32_t l = __ldg(ids + blockIdx.x);
	reinterpret_cast<float4*>(dst + uint64_t(blockIdx.x) * 512u)[threadIdx.x] = __ldg(reinterpret_cast<const float4*>(field + uint64_t(l) * 512u) + threadIdx.x);
}
__global__ void __launch_bounds__(128) k_unpack_leaves(float* __restrict__ field, const int32_t* __restrict__ ids, const float* __restrict__ src) {
	const int32_t l = __ldg(ids + blockIdx.x);
	reinterpret_cast<float4*>(field + uint64_t(l) * 512u)[threadIdx.x] = __ldg(reinterpret_cast<const float4*>(src + uint64_t(blockIdx.x) * 512u) + threadIdx.x);
}
void launch_pack_leaves(const float* field, const int32_t* ids, uint64_t n_ids, float* dst, cudaStream_t st) {
	if (n_ids) HNS_LAUNCH(k_pack_leaves, unsigned(n_ids), 128, 0, st, field, ids, dst);
}
void launch_unpack_leaves(float* field, const int32_t* ids, uint64_t n_ids, const float* src, cudaStream_t st) {
	if (n_ids) HNS_LAUNCH(k_unpack_leaves, unsigned(n_ids), 128, 0, st, field, ids, src);
}

}  // namespace hns
