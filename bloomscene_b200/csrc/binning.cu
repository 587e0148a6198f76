// Binning for sm_100a: depth sort, two-level tile binning, tile ranges.
//
// The reference builds 64-bit keys (tile << 32 | depth bits) for all R (tile, Gaussian) instances
// and runs one 6-pass CUB radix sort over 12-byte pairs (rasterizer_impl.cu:70-111, 278-318).
// The same order — ascending (tile, depth bits, Gaussian id), which is what a stable sort of the
// reference's emission order yields — is produced here without ever sorting R-sized arrays:
//
//   1. stable sort of the P Gaussians by depth bits (u32 key, value = id; culled ones carry the
//      key 0xFFFFFFFF, emit nothing, and may land anywhere)                    -> `order`
//   2. COARSE level: one fused kernel scans supertiles-per-Gaussian in depth order (decoupled
//      look-back, 32 predecessors per poll) and emits (supertile id, Gaussian id) instances, a
//      supertile being 8x8 tiles; R1 ~ 1.3 V of them.  One stable radix pass on the supertile id
//      (<= 8 bits up to 2048x2048 pixels) turns that into per-supertile lists in depth order.
//   3. FINE level: every supertile list is cut into slices of 128 entries, one warp per slice.
//      count  : per slice, how many entries cover each of the supertile's 64 tiles (lane = tile owner,
//               entries broadcast one by one in list order)
//      scan   : per (supertile, tile) exclusive prefix over the slices; then one exclusive scan over
//               all tile ids gives every tile's start in `point_list` and the tile `ranges`
//               (reference identifyTileRanges, rasterizer_impl.cu:116-138: empty tiles stay (0,0))
//      scatter: the count loop again; every tile owner appends the ids covering its tile at its running
//               position, so ids are written straight to their final place in stable order.
//
// Because (1) is stable in id, (2) is stable, and slices / lanes are walked in list order, instances
// inside a tile end up ordered by (depth bits, id): exactly the reference's sorted `point_list`.
// Traffic is ~70 P + 16 R1 + 4 R bytes instead of the reference's 152 R.
//
// The radix sort is hand-written.  Each pass over one digit is three kernels:
//   upsweep   : per 2048-key tile, digit counts (warp match_any + shared atomics)  -> table[digit][tile]
//   scan      : one block per digit, exclusive scan along the tiles + digit totals
//   downsweep : per tile, stable ranks (match_any + shared atomics returning the old value, so the
//               16 keys of a thread are in flight together), scatter through shared memory so
//               global writes are coalesced per digit run.
// A single-pass chained scan ("onesweep") was measured first: with ~600 tiles resident at once the
// first wave spends ~100 us per pass polling unpublished predecessors on this GPU, far more than
// the extra 4 bytes/key the upsweep reads (see profiles/).
#include "common.cuh"
#include "kernels.h"

namespace brs {

namespace {

constexpr int SORT_THREADS = 256;
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int SORT_ITEMS = 8;
constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS; // 2048 pairs per tile (smaller tiles = more CTAs in flight: the passes are latency-bound chains of MATCH/ATOMS)
constexpr int RADIX_MAX = 256;
constexpr int MAX_PASSES = 4;

constexpr uint32_t FLAG_LOCAL = 1u << 30; // (emit scan) block-local count published
constexpr uint32_t FLAG_INCL = 2u << 30;  // (emit scan) inclusive prefix published
constexpr uint32_t FLAG_MASK = 3u << 30;
constexpr uint32_t VALUE_MASK = ~FLAG_MASK;

__device__ __forceinline__ uint32_t ld_relaxed(const uint32_t* p)
{
	uint32_t v;
	asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ void st_relaxed(uint32_t* p, uint32_t v)
{
	asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Exclusive scan of one value per thread over a 256-thread block; also returns the block total.
// s_warp must hold SORT_WARPS uint32.  Contains two __syncthreads.
__device__ __forceinline__ uint32_t block_exclusive_scan_256(uint32_t v, uint32_t* s_warp, uint32_t& total)
{
	const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	uint32_t incl = v;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		uint32_t n = __shfl_up_sync(0xffffffffu, incl, o);
		if (lane >= (uint32_t)o)
			incl += n;
	}
	if (lane == 31)
		s_warp[warp] = incl;
	__syncthreads();
	uint32_t warp_off = 0, tot = 0;
#pragma unroll
	for (int w = 0; w < SORT_WARPS; w++) {
		const uint32_t c = s_warp[w];
		if ((uint32_t)w < warp)
			warp_off += c;
		tot += c;
	}
	__syncthreads();
	total = tot;
	return warp_off + incl - v;
}

// ---- upsweep: digit counts of one tile ------------------------------------------------------------
__global__ void __launch_bounds__(SORT_THREADS)
    upsweep_kernel(const uint32_t* __restrict__ keys, uint32_t n, uint32_t bias, int shift, int bits, uint32_t tiles,
                   uint32_t* __restrict__ table, uint32_t drop_key, int drop)
{
	__shared__ uint32_t s_cnt[RADIX_MAX];
	const uint32_t tid = threadIdx.x, lane = tid & 31;
	const uint32_t radix = 1u << bits, mask = radix - 1u;
	s_cnt[tid] = 0;
	__syncthreads();
	const uint32_t tile = blockIdx.x;
	const uint32_t base = tile * SORT_TILE;
	const uint32_t valid_count = min((uint32_t)SORT_TILE, n - base);
	uint32_t key[SORT_ITEMS];
#pragma unroll
	for (int i = 0; i < SORT_ITEMS; i++) {
		const uint32_t li = tid + i * SORT_THREADS;
		key[i] = (li < valid_count) ? __ldg(keys + base + li) : 0xffffffffu;
	}
#pragma unroll
	for (int i = 0; i < SORT_ITEMS; i++) {
		const bool valid = (tid + i * SORT_THREADS) < valid_count && !(drop && key[i] == drop_key);
		const uint32_t d = ((key[i] - bias) >> shift) & mask;
		const uint32_t peers = __match_any_sync(0xffffffffu, valid ? d : 0xffffffffu);
		if (valid && (uint32_t)(__ffs(peers) - 1) == lane)
			atomicAdd(s_cnt + d, (uint32_t)__popc(peers));
	}
	__syncthreads();
	if (tid < radix)
		table[(size_t)tid * tiles + tile] = s_cnt[tid];
}

// ---- scan: per digit, exclusive prefix over the tiles (in place) + digit total --------------------
__global__ void __launch_bounds__(SORT_THREADS) scan_kernel(uint32_t* __restrict__ table, uint32_t tiles,
                                                           uint32_t* __restrict__ totals)
{
	__shared__ uint32_t s_warp[SORT_WARPS];
	uint32_t* row = table + (size_t)blockIdx.x * tiles;
	uint32_t carry = 0;
	for (uint32_t t0 = 0; t0 < tiles; t0 += SORT_THREADS) {
		const uint32_t t = t0 + threadIdx.x;
		const uint32_t v = (t < tiles) ? row[t] : 0u;
		uint32_t total;
		const uint32_t ex = block_exclusive_scan_256(v, s_warp, total);
		if (t < tiles)
			row[t] = carry + ex;
		carry += total;
	}
	if (threadIdx.x == 0)
		totals[blockIdx.x] = carry;
}

// ---- downsweep: stable scatter of one tile --------------------------------------------------------
__global__ void __launch_bounds__(SORT_THREADS)
    downsweep_kernel(const uint32_t* __restrict__ kin, const uint32_t* __restrict__ vin, uint32_t* __restrict__ kout,
                     uint32_t* __restrict__ vout, uint32_t n, uint32_t bias, int shift, int bits, uint32_t tiles,
                     const uint32_t* __restrict__ table, const uint32_t* __restrict__ totals, uint32_t drop_key, int drop)
{
	__shared__ uint32_t s_cnt[SORT_WARPS][RADIX_MAX];
	__shared__ uint32_t s_keys[SORT_TILE];
	__shared__ uint32_t s_vals[SORT_TILE];
	__shared__ uint32_t s_digit_start[RADIX_MAX];
	__shared__ uint32_t s_goff[RADIX_MAX];
	__shared__ uint32_t s_warp[SORT_WARPS];

	const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const uint32_t radix = 1u << bits, mask = radix - 1u;
	for (int i = tid; i < SORT_WARPS * RADIX_MAX; i += SORT_THREADS)
		(&s_cnt[0][0])[i] = 0;
	__syncthreads();
	const uint32_t tile = blockIdx.x;
	const uint32_t base = tile * SORT_TILE;
	const uint32_t valid_count = min((uint32_t)SORT_TILE, n - base);

	// load keys (warp-striped: coalesced, and rank order == index order)
	uint32_t key[SORT_ITEMS], rank[SORT_ITEMS];
	const uint32_t wbase = warp * (32 * SORT_ITEMS) + lane;
#pragma unroll
	for (int i = 0; i < SORT_ITEMS; i++) {
		const uint32_t li = wbase + i * 32;
		key[i] = (li < valid_count) ? __ldg(kin + base + li) : 0xffffffffu;
	}
	// global offsets of this tile's digits (independent of the ranking below)
	const uint32_t tile_excl = (tid < radix) ? __ldg(table + (size_t)tid * tiles + tile) : 0u;
	const uint32_t digit_total = (tid < radix) ? __ldg(totals + tid) : 0u;

	// rank inside the warp (stable).  match_any groups equal digits; the group's first lane bumps the
	// warp-private counter with a shared-memory atomic that returns the old value.  Same-address
	// atomics of one warp complete in program order, so no barrier is needed and all SORT_ITEMS
	// chains (MATCH -> ATOMS -> SHFL) overlap.
	uint32_t* my_cnt = s_cnt[warp];
#pragma unroll
	for (int i = 0; i < SORT_ITEMS; i++) {
		const bool valid = (wbase + i * 32) < valid_count && !(drop && key[i] == drop_key);
		const uint32_t d = ((key[i] - bias) >> shift) & mask;
		const uint32_t peers = __match_any_sync(0xffffffffu, valid ? d : 0xffffffffu);
		const int leader = __ffs(peers) - 1;
		uint32_t old = 0;
		if (valid && (int)lane == leader)
			old = atomicAdd(my_cnt + d, (uint32_t)__popc(peers));
		old = __shfl_sync(0xffffffffu, old, leader);
		rank[i] = old + __popc(peers & lanemask_lt());
	}
	__syncthreads();

	// per digit: exclusive offsets of the warps, tile total
	uint32_t total = 0;
	if (tid < radix) {
#pragma unroll
		for (int w = 0; w < SORT_WARPS; w++) {
			const uint32_t c = s_cnt[w][tid];
			s_cnt[w][tid] = total;
			total += c;
		}
	}
	uint32_t dummy, kept; // kept = keys of this tile that take part (all valid ones unless `drop`)
	const uint32_t digit_start = block_exclusive_scan_256(total, s_warp, kept);
	const uint32_t bin_base = block_exclusive_scan_256(digit_total, s_warp, dummy);
	if (tid < radix) {
		s_digit_start[tid] = digit_start;
		s_goff[tid] = bin_base + tile_excl - digit_start; // global position = s_goff[d] + position in tile
	}
	__syncthreads();

	// scatter into shared memory in sorted-by-digit order (values are fetched only now)
#pragma unroll
	for (int i = 0; i < SORT_ITEMS; i++) {
		const uint32_t li = wbase + i * 32;
		if (li < valid_count && !(drop && key[i] == drop_key)) {
			const uint32_t d = ((key[i] - bias) >> shift) & mask;
			const uint32_t pos = s_digit_start[d] + s_cnt[warp][d] + rank[i];
			s_keys[pos] = key[i];
			s_vals[pos] = vin ? __ldg(vin + base + li) : base + li;
		}
	}
	__syncthreads();

	// coalesced write-out: consecutive threads hold consecutive positions of a digit run
#pragma unroll
	for (int i = 0; i < SORT_ITEMS; i++) {
		const uint32_t j = tid + i * SORT_THREADS;
		if (j < kept) {
			const uint32_t k = s_keys[j];
			const uint32_t g = s_goff[((k - bias) >> shift) & mask] + j;
			kout[g] = k;
			vout[g] = s_vals[j];
		}
	}
}

struct SortScratch {
	uint32_t* table;  // [RADIX_MAX][tiles] (reused by every pass)
	uint32_t* totals; // [RADIX_MAX]
	uint32_t* tmp_keys;
	uint32_t* tmp_vals;
	size_t total_bytes;
};

SortScratch carve_sort_scratch(void* scratch, size_t n)
{
	SortScratch s{};
	const size_t tiles = (n + SORT_TILE - 1) / SORT_TILE;
	char* p = static_cast<char*>(scratch);
	size_t off = 0;
	s.table = reinterpret_cast<uint32_t*>(p + off);
	off += align_up(sizeof(uint32_t) * RADIX_MAX * (tiles ? tiles : 1), 256);
	s.totals = reinterpret_cast<uint32_t*>(p + off);
	off += align_up(sizeof(uint32_t) * RADIX_MAX, 256);
	s.tmp_keys = reinterpret_cast<uint32_t*>(p + off);
	off += align_up(sizeof(uint32_t) * n, 256);
	s.tmp_vals = reinterpret_cast<uint32_t*>(p + off);
	off += align_up(sizeof(uint32_t) * n, 256);
	s.total_bytes = off;
	return s;
}

} // namespace

size_t sort_scratch_bytes(size_t n) { return carve_sort_scratch(nullptr, n).total_bytes; }

void sort_tmp_buffers(void* scratch, size_t n, uint32_t** tmp_keys, uint32_t** tmp_vals)
{
	SortScratch s = carve_sort_scratch(scratch, n);
	*tmp_keys = s.tmp_keys;
	*tmp_vals = s.tmp_vals;
}

cudaError_t sort_pass(const uint32_t* kin, const uint32_t* vin, uint32_t* kout, uint32_t* vout, size_t n, uint32_t bias,
                      int shift, int bits, void* scratch, cudaStream_t stream, bool drop, uint32_t drop_key)
{
	if (n == 0)
		return cudaSuccess;
	SortScratch s = carve_sort_scratch(scratch, n);
	const uint32_t tiles = (uint32_t)((n + SORT_TILE - 1) / SORT_TILE);
	upsweep_kernel<<<tiles, SORT_THREADS, 0, stream>>>(kin, (uint32_t)n, bias, shift, bits, tiles, s.table, drop_key,
	                                                   drop ? 1 : 0);
	scan_kernel<<<1u << bits, SORT_THREADS, 0, stream>>>(s.table, tiles, s.totals);
	downsweep_kernel<<<tiles, SORT_THREADS, 0, stream>>>(kin, vin, kout, vout, (uint32_t)n, bias, shift, bits, tiles,
	                                                     s.table, s.totals, drop_key, drop ? 1 : 0);
	count_launch();
	count_launch();
	count_launch();
	return cudaGetLastError();
}

cudaError_t sort_pairs(const uint32_t* keys_in, const uint32_t* vals_in, uint32_t* keys_out, uint32_t* vals_out,
                       size_t n, int begin_bit, int end_bit, void* scratch, cudaStream_t stream)
{
	if (n == 0)
		return cudaSuccess;
	int nbits = end_bit - begin_bit;
	if (nbits < 1)
		nbits = 1;
	const int passes = (nbits + 7) / 8;
	const int base_bits = nbits / passes, extra = nbits % passes;
	SortScratch s = carve_sort_scratch(scratch, n);
	const uint32_t* kin = keys_in;
	const uint32_t* vin = vals_in;
	int shift = begin_bit;
	for (int p = 0; p < passes; p++) {
		const int bits = base_bits + (p < extra ? 1 : 0);
		const bool to_out = ((passes - 1 - p) & 1) == 0;
		uint32_t* ko = to_out ? keys_out : s.tmp_keys;
		uint32_t* vo = to_out ? vals_out : s.tmp_vals;
		cudaError_t e = sort_pass(kin, vin, ko, vo, n, 0u, shift, bits, scratch, stream);
		if (e != cudaSuccess)
			return e;
		kin = ko;
		vin = vo;
		shift += bits;
	}
	return cudaGetLastError();
}

// ---- fused scan + emission ------------------------------------------------------------------------

namespace {

constexpr int EMIT_THREADS = 256;
constexpr int EMIT_ITEMS = 4;
constexpr int EMIT_TILE = EMIT_THREADS * EMIT_ITEMS; // Gaussians per block

// Tile rectangle -> rectangle in cells of (1 << shift) tiles; an empty rectangle stays empty.
__device__ __forceinline__ uint2 coarsen_rect(uint2 r, uint32_t shift)
{
	const uint32_t x0 = r.x & 0xffffu, x1 = r.x >> 16, y0 = r.y & 0xffffu, y1 = r.y >> 16;
	if (x1 <= x0 || y1 <= y0)
		return make_uint2(0u, 0u);
	return make_uint2((x0 >> shift) | ((((x1 - 1) >> shift) + 1) << 16), (y0 >> shift) | ((((y1 - 1) >> shift) + 1) << 16));
}

__global__ void __launch_bounds__(EMIT_THREADS)
    emit_kernel(const uint32_t* __restrict__ order, const uint2* __restrict__ rect, uint32_t P, uint32_t shift,
                uint32_t grid_x, uint32_t* __restrict__ tile_keys, uint32_t* __restrict__ inst_ids, uint32_t R,
                uint32_t* status, uint32_t* ticket)
{
	__shared__ uint32_t s_incl[EMIT_TILE]; // inclusive instance offsets inside this block
	__shared__ uint32_t s_id[EMIT_TILE];
	__shared__ uint2 s_rect[EMIT_TILE];
	__shared__ uint32_t s_warp[SORT_WARPS];
	__shared__ uint32_t s_bcast[2];

	const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	if (tid == 0)
		s_bcast[0] = atomicAdd(ticket, 1u);
	__syncthreads();
	const uint32_t blk = s_bcast[0];
	const uint32_t first = blk * EMIT_TILE + tid * EMIT_ITEMS;

	uint32_t cnt[EMIT_ITEMS];
	uint32_t tsum = 0;
#pragma unroll
	for (int k = 0; k < EMIT_ITEMS; k++) {
		const uint32_t s = first + k;
		uint32_t id = 0;
		uint2 r = make_uint2(0u, 0u);
		if (s < P) {
			id = __ldg(order + s);
			r = coarsen_rect(__ldg(rect + id), shift);
		}
		const uint32_t w = (r.x >> 16) - (r.x & 0xffffu);
		const uint32_t hgt = (r.y >> 16) - (r.y & 0xffffu);
		cnt[k] = w * hgt;
		tsum += cnt[k];
		s_id[tid * EMIT_ITEMS + k] = id;
		s_rect[tid * EMIT_ITEMS + k] = r;
	}
	uint32_t block_total;
	uint32_t run = block_exclusive_scan_256(tsum, s_warp, block_total);
#pragma unroll
	for (int k = 0; k < EMIT_ITEMS; k++) {
		run += cnt[k];
		s_incl[tid * EMIT_ITEMS + k] = run;
	}

	// chained scan over blocks (single value): decoupled look-back by the whole first warp,
	// 32 predecessors per poll
	if (warp == 0) {
		if (lane == 0)
			st_relaxed(status + blk, (blk == 0 ? FLAG_INCL : FLAG_LOCAL) | block_total);
		uint32_t excl = 0;
		int p = (int)blk - 1;
		while (p >= 0) {
			const int q = p - (int)lane;
			const uint32_t v = (q >= 0) ? ld_relaxed(status + q) : FLAG_INCL;
			const uint32_t f = v & FLAG_MASK;
			const uint32_t not_ready = __ballot_sync(0xffffffffu, f == 0);
			const uint32_t inclusive = __ballot_sync(0xffffffffu, f == FLAG_INCL);
			const int first_nr = not_ready ? __ffs(not_ready) - 1 : 32;
			const int first_in = inclusive ? __ffs(inclusive) - 1 : 32;
			const int take = (first_in < first_nr) ? first_in + 1 : first_nr; // lanes [0, take) are summed
			uint32_t c = ((int)lane < take) ? (v & VALUE_MASK) : 0u;
#pragma unroll
			for (int o = 16; o > 0; o >>= 1)
				c += __shfl_xor_sync(0xffffffffu, c, o);
			excl += c;
			if (first_in < first_nr)
				break;
			p -= first_nr;
		}
		if (lane == 0) {
			if (blk > 0)
				st_relaxed(status + blk, FLAG_INCL | (excl + block_total));
			s_bcast[1] = excl;
		}
	}
	__syncthreads();
	const uint32_t gbase = s_bcast[1];

	// balanced emission: one output slot per thread-iteration, owner found by binary search
	for (uint32_t j = tid; j < block_total; j += EMIT_THREADS) {
		int lo = 0, hi = EMIT_TILE - 1; // first index with s_incl > j
		while (lo < hi) {
			const int mid = (lo + hi) >> 1;
			if (s_incl[mid] > j)
				hi = mid;
			else
				lo = mid + 1;
		}
		const uint2 r = s_rect[lo];
		const uint32_t x0 = r.x & 0xffffu, w = (r.x >> 16) - x0, y0 = r.y & 0xffffu;
		const uint32_t hgt = (r.y >> 16) - y0;
		const uint32_t k = j - (s_incl[lo] - w * hgt);
		const uint32_t row = k / w;
		const uint32_t g = gbase + j;
		if (g < R) {
			tile_keys[g] = (y0 + row) * grid_x + x0 + (k - row * w);
			inst_ids[g] = s_id[lo];
		}
	}
}

__global__ void __launch_bounds__(256)
    tile_ranges_kernel(const uint32_t* __restrict__ keys, uint32_t R, uint2* __restrict__ ranges)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= R)
		return;
	const uint32_t cur = __ldg(keys + i);
	if (i == 0)
		ranges[cur].x = 0;
	else {
		const uint32_t prev = __ldg(keys + i - 1);
		if (cur != prev) {
			ranges[prev].y = i;
			ranges[cur].x = i;
		}
	}
	if (i == R - 1)
		ranges[cur].y = R;
}

} // namespace

size_t emit_scratch_bytes(size_t P)
{
	const size_t blocks = (P + EMIT_TILE - 1) / EMIT_TILE;
	return align_up(256 + sizeof(uint32_t) * blocks, 256);
}

cudaError_t launch_emit(const uint32_t* order, const uint2* rect, size_t P, uint32_t shift, uint32_t grid_x,
                        uint32_t* tile_keys, uint32_t* inst_ids, size_t R, void* scratch, cudaStream_t stream)
{
	if (P == 0)
		return cudaSuccess;
	const uint32_t blocks = (uint32_t)((P + EMIT_TILE - 1) / EMIT_TILE);
	cudaError_t e = cudaMemsetAsync(scratch, 0, emit_scratch_bytes(P), stream);
	if (e != cudaSuccess)
		return e;
	uint32_t* ticket = static_cast<uint32_t*>(scratch);
	uint32_t* status = ticket + 64;
	emit_kernel<<<blocks, EMIT_THREADS, 0, stream>>>(order, rect, (uint32_t)P, shift, grid_x, tile_keys, inst_ids,
	                                                 (uint32_t)R, status, ticket);
	count_launch();
	return cudaGetLastError();
}

// ---- fine level: supertile lists -> per-tile lists --------------------------------------------------

namespace {

constexpr int FINE_WARPS = 8; // warps (= slices) per CTA of the count / scatter kernels

__device__ __forceinline__ uint32_t spread4(uint32_t b) // bit i of b (i < 4) -> bit 8 i
{
	return (b * 0x00204081u) & 0x01010101u;
}

// Exclusive scan of ceil(n_s / FINE_SLICE) over the supertiles -> slice_base[0..ns]; also copies the
// list starts.  One CTA (ns is a few hundred).
__global__ void __launch_bounds__(SORT_THREADS)
    fine_plan_kernel(const uint2* __restrict__ coarse_ranges, uint32_t ns, uint32_t* __restrict__ slice_base)
{
	__shared__ uint32_t s_warp[SORT_WARPS];
	uint32_t carry = 0;
	for (uint32_t s0 = 0; s0 < ns; s0 += SORT_THREADS) {
		const uint32_t s = s0 + threadIdx.x;
		uint32_t v = 0;
		if (s < ns) {
			const uint2 r = coarse_ranges[s];
			v = (r.y - r.x + FINE_SLICE - 1) / FINE_SLICE;
		}
		uint32_t total;
		const uint32_t ex = block_exclusive_scan_256(v, s_warp, total);
		if (s < ns)
			slice_base[s] = carry + ex;
		carry += total;
	}
	if (threadIdx.x == 0)
		slice_base[ns] = carry;
}

// One warp per slice of <= FINE_SLICE consecutive entries of one supertile's list.  Entries are read 32
// at a time (lane = entry, list order): each lane builds the 64-bit mask of the supertile's tiles its
// rectangle covers.  Then the roles flip: lane t OWNS local tiles t (rows 0..3) and t + 32 (rows 4..7)
// and the warp walks the 32 entries in list order, broadcasting one entry's mask (and id) per step.
// The owner of a covered tile bumps its private running position (COUNT pass) and writes the id there
// (SCATTER pass) - no ballots, no ranks: walking entries in order IS the stable order.
template <bool SCATTER>
__global__ void __launch_bounds__(FINE_WARPS * 32)
    fine_kernel(const uint32_t* __restrict__ coarse_list, const uint2* __restrict__ coarse_ranges,
                const uint32_t* __restrict__ slice_base, const uint2* __restrict__ rect, uint32_t ns, uint32_t ns_x,
                uint32_t grid_x, uint32_t grid_y, uint32_t* __restrict__ table, const uint32_t* __restrict__ tile_start,
                uint32_t* __restrict__ point_list)
{
	const uint32_t lane = threadIdx.x & 31;
	const uint32_t sl = blockIdx.x * FINE_WARPS + (threadIdx.x >> 5);
	if (sl >= __ldg(slice_base + ns))
		return;
	// supertile of this slice: last s with slice_base[s] <= sl
	uint32_t lo_s = 0, hi_s = ns;
	while (hi_s - lo_s > 1) {
		const uint32_t mid = (lo_s + hi_s) >> 1;
		if (__ldg(slice_base + mid) <= sl)
			lo_s = mid;
		else
			hi_s = mid;
	}
	const uint32_t st = lo_s;
	const uint2 cr = __ldg(coarse_ranges + st);
	const uint32_t begin = cr.x + (sl - __ldg(slice_base + st)) * FINE_SLICE;
	const uint32_t end = min(begin + (uint32_t)FINE_SLICE, cr.y);
	const uint32_t tx0 = (st % ns_x) << ST_SHIFT, ty0 = (st / ns_x) << ST_SHIFT;

	// lane l keeps the running value of local tiles l (rows 0..3) and l + 32 (rows 4..7)
	uint32_t run0 = 0, run1 = 0;
	if (SCATTER) {
		const uint32_t x = tx0 + (lane & 7), y = ty0 + (lane >> 3);
		if (x < grid_x && y < grid_y)
			run0 = __ldg(tile_start + y * grid_x + x) + table[(size_t)sl * ST_TILES + lane];
		if (x < grid_x && y + 4 < grid_y)
			run1 = __ldg(tile_start + (y + 4) * grid_x + x) + table[(size_t)sl * ST_TILES + 32 + lane];
	}

	for (uint32_t e0 = begin; e0 < end; e0 += 32) {
		const uint32_t e = e0 + lane;
		uint32_t id = 0, lo = 0, hi = 0;
		if (e < end) {
			id = __ldg(coarse_list + e);
			const uint2 r = __ldg(rect + id);
			const int x0 = (int)(r.x & 0xffffu) - (int)tx0, x1 = (int)(r.x >> 16) - (int)tx0;
			const int y0 = (int)(r.y & 0xffffu) - (int)ty0, y1 = (int)(r.y >> 16) - (int)ty0;
			const uint32_t cx0 = (uint32_t)max(x0, 0), cx1 = (uint32_t)min(x1, 8);
			const uint32_t cy0 = (uint32_t)max(y0, 0), cy1 = (uint32_t)min(y1, 8);
			if (cx1 > cx0 && cy1 > cy0) {
				const uint32_t cols = ((1u << cx1) - (1u << cx0)) & 0xffu;
				const uint32_t rows = ((1u << cy1) - (1u << cy0)) & 0xffu;
				lo = cols * spread4(rows & 15u);
				hi = cols * spread4(rows >> 4);
			}
		}
		const uint32_t cnt = min(32u, end - e0);
		if (SCATTER) {
#pragma unroll 8
			for (uint32_t j = 0; j < cnt; j++) {
				const uint32_t mlo = __shfl_sync(0xffffffffu, lo, j);
				const uint32_t mhi = __shfl_sync(0xffffffffu, hi, j);
				const uint32_t idj = __shfl_sync(0xffffffffu, id, j);
				if ((mlo >> lane) & 1u)
					point_list[run0++] = idj;
				if ((mhi >> lane) & 1u)
					point_list[run1++] = idj;
			}
		} else {
#pragma unroll 8
			for (uint32_t j = 0; j < cnt; j++) {
				const uint32_t mlo = __shfl_sync(0xffffffffu, lo, j);
				const uint32_t mhi = __shfl_sync(0xffffffffu, hi, j);
				run0 += (mlo >> lane) & 1u;
				run1 += (mhi >> lane) & 1u;
			}
		}
	}
	if (!SCATTER) {
		table[(size_t)sl * ST_TILES + lane] = run0;
		table[(size_t)sl * ST_TILES + 32 + lane] = run1;
	}
}

// Per supertile (one CTA, 4 groups x 64 tiles): exclusive prefix of the slice counts along the
// slices, in place, and the per-tile totals into tile_count[tile id].
__global__ void __launch_bounds__(256)
    fine_scan_kernel(uint32_t* __restrict__ table, const uint32_t* __restrict__ slice_base, uint32_t ns_x,
                     uint32_t grid_x, uint32_t grid_y, uint32_t* __restrict__ tile_count)
{
	__shared__ uint32_t s_part[4][ST_TILES];
	const uint32_t st = blockIdx.x;
	const uint32_t first = __ldg(slice_base + st), nsl = __ldg(slice_base + st + 1) - first;
	const uint32_t t = threadIdx.x & 63, q = threadIdx.x >> 6;
	const uint32_t per = (nsl + 3) / 4;
	const uint32_t j0 = min(q * per, nsl), j1 = min(j0 + per, nsl);
	uint32_t* col = table + (size_t)first * ST_TILES + t;
	uint32_t sum = 0;
#pragma unroll 8
	for (uint32_t j = j0; j < j1; j++)
		sum += col[(size_t)j * ST_TILES];
	s_part[q][t] = sum;
	__syncthreads();
	uint32_t run = 0, total = 0;
#pragma unroll
	for (uint32_t k = 0; k < 4; k++) {
		const uint32_t v = s_part[k][t];
		if (k < q)
			run += v;
		total += v;
	}
#pragma unroll 8
	for (uint32_t j = j0; j < j1; j++) {
		const uint32_t v = col[(size_t)j * ST_TILES];
		col[(size_t)j * ST_TILES] = run;
		run += v;
	}
	if (q == 0) {
		const uint32_t x = ((st % ns_x) << ST_SHIFT) + (t & 7), y = ((st / ns_x) << ST_SHIFT) + (t >> 3);
		if (x < grid_x && y < grid_y)
			tile_count[y * grid_x + x] = total;
	}
}

// Exclusive scan of the per-tile counts in tile-id order -> tile_start, and the tile ranges
// (reference identifyTileRanges: tiles without instances keep (0, 0)).  One CTA.
__global__ void __launch_bounds__(1024)
    tile_offsets_kernel(const uint32_t* __restrict__ tile_count, uint32_t num_tiles, uint32_t* __restrict__ tile_start,
                        uint2* __restrict__ ranges)
{
	__shared__ uint32_t s_warp[32];
	__shared__ uint32_t s_carry;
	const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	if (threadIdx.x == 0)
		s_carry = 0;
	__syncthreads();
	for (uint32_t t0 = 0; t0 < num_tiles; t0 += 1024) {
		const uint32_t t = t0 + threadIdx.x;
		const uint32_t v = (t < num_tiles) ? tile_count[t] : 0u;
		uint32_t incl = v;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			const uint32_t n = __shfl_up_sync(0xffffffffu, incl, o);
			if (lane >= (uint32_t)o)
				incl += n;
		}
		if (lane == 31)
			s_warp[warp] = incl;
		__syncthreads();
		if (warp == 0) {
			uint32_t w = s_warp[lane], wi = w;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1) {
				const uint32_t n = __shfl_up_sync(0xffffffffu, wi, o);
				if (lane >= (uint32_t)o)
					wi += n;
			}
			s_warp[lane] = wi - w; // exclusive over warps
		}
		__syncthreads();
		const uint32_t start = s_carry + s_warp[warp] + incl - v;
		if (t < num_tiles) {
			tile_start[t] = start;
			ranges[t] = v ? make_uint2(start, start + v) : make_uint2(0u, 0u);
		}
		__syncthreads();
		if (threadIdx.x == 1023)
			s_carry = start + v;
		__syncthreads();
	}
}

struct FineScratch {
	uint2* coarse_ranges;  // [ns]
	uint32_t* slice_base;  // [ns + 1]
	uint32_t* tile_count;  // [num_tiles]
	uint32_t* tile_start;  // [num_tiles]
	uint32_t* table;       // [max_slices][64]
	size_t total_bytes;
};

FineScratch carve_fine_scratch(void* scratch, size_t R1, uint32_t grid_x, uint32_t grid_y)
{
	FineScratch f{};
	const size_t ns = (size_t)supertiles(grid_x) * supertiles(grid_y);
	const size_t num_tiles = (size_t)grid_x * grid_y;
	const size_t max_slices = R1 / FINE_SLICE + ns;
	char* p = static_cast<char*>(scratch);
	size_t off = 0;
	f.coarse_ranges = reinterpret_cast<uint2*>(p + off);
	off += align_up(sizeof(uint2) * ns, 256);
	f.slice_base = reinterpret_cast<uint32_t*>(p + off);
	off += align_up(sizeof(uint32_t) * (ns + 1), 256);
	f.tile_count = reinterpret_cast<uint32_t*>(p + off);
	off += align_up(sizeof(uint32_t) * num_tiles, 256);
	f.tile_start = reinterpret_cast<uint32_t*>(p + off);
	off += align_up(sizeof(uint32_t) * num_tiles, 256);
	f.table = reinterpret_cast<uint32_t*>(p + off);
	off += align_up(sizeof(uint32_t) * ST_TILES * (max_slices ? max_slices : 1), 256);
	f.total_bytes = off;
	return f;
}

} // namespace

size_t fine_scratch_bytes(size_t R1, uint32_t grid_x, uint32_t grid_y)
{
	return carve_fine_scratch(nullptr, R1, grid_x, grid_y).total_bytes;
}

cudaError_t launch_fine_binning(const uint32_t* sorted_coarse_keys, const uint32_t* coarse_list, size_t R1,
                                const uint2* rect, uint32_t grid_x, uint32_t grid_y, uint32_t* point_list,
                                uint2* ranges, void* scratch, cudaStream_t stream)
{
	const uint32_t ns_x = supertiles(grid_x), ns = ns_x * supertiles(grid_y);
	const uint32_t num_tiles = grid_x * grid_y;
	if (num_tiles == 0)
		return cudaSuccess;
	FineScratch f = carve_fine_scratch(scratch, R1, grid_x, grid_y);
	// coarse_ranges and tile_count start at zero (supertiles / tiles nothing touches are never written)
	cudaError_t e = cudaMemsetAsync(f.coarse_ranges, 0, (char*)f.tile_start - (char*)f.coarse_ranges, stream);
	if (e != cudaSuccess)
		return e;
	if (R1 > 0) {
		tile_ranges_kernel<<<(uint32_t)((R1 + 255) / 256), 256, 0, stream>>>(sorted_coarse_keys, (uint32_t)R1,
		                                                                      f.coarse_ranges);
		count_launch();
	}
	fine_plan_kernel<<<1, SORT_THREADS, 0, stream>>>(f.coarse_ranges, ns, f.slice_base);
	count_launch();
	const uint32_t max_slices = (uint32_t)(R1 / FINE_SLICE + ns);
	const uint32_t blocks = (max_slices + FINE_WARPS - 1) / FINE_WARPS;
	if (R1 > 0) {
		fine_kernel<false><<<blocks, FINE_WARPS * 32, 0, stream>>>(coarse_list, f.coarse_ranges, f.slice_base, rect, ns,
		                                                            ns_x, grid_x, grid_y, f.table, nullptr, nullptr);
		count_launch();
		fine_scan_kernel<<<ns, 256, 0, stream>>>(f.table, f.slice_base, ns_x, grid_x, grid_y, f.tile_count);
		count_launch();
	}
	tile_offsets_kernel<<<1, 1024, 0, stream>>>(f.tile_count, num_tiles, f.tile_start, ranges);
	count_launch();
	if (R1 > 0) {
		fine_kernel<true><<<blocks, FINE_WARPS * 32, 0, stream>>>(coarse_list, f.coarse_ranges, f.slice_base, rect, ns,
		                                                           ns_x, grid_x, grid_y, f.table, f.tile_start, point_list);
		count_launch();
	}
	return cudaGetLastError();
}

cudaError_t launch_tile_ranges(const uint32_t* sorted_tile_keys, size_t R, uint2* ranges, cudaStream_t stream)
{
	if (R == 0)
		return cudaSuccess;
	tile_ranges_kernel<<<(uint32_t)((R + 255) / 256), 256, 0, stream>>>(sorted_tile_keys, (uint32_t)R, ranges);
	count_launch();
	return cudaGetLastError();
}

} // namespace brs
