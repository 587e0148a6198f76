// Binning for sm_100a: depth sort, two-level tile binning, tile ranges.
//
// The reference builds 64-bit keys (tile << 32 | depth bits) for all R (tile, Gaussian) instances
// and runs one 6-pass CUB radix sort over 12-byte pairs (rasterizer_impl.cu:70-111, 278-318).
// The same order — ascending (tile, depth bits, Gaussian id), which is what a stable sort of the
// reference's emission order yields — is produced here without ever sorting R-sized arrays:
//
//   1. stable sort of the P Gaussians by depth bits (u32 key, value = id; culled ones carry the
//      key 0xFFFFFFFF and are dropped by the first pass)                               -> `order`
//   2. COARSE level: one fused kernel scans supertiles-per-Gaussian in depth order (decoupled
//      look-back, 32 predecessors per poll), emits (supertile id, Gaussian id) instances — a
//      supertile being 8x8 tiles; R1 ~ 1.3 V of them — and histograms the supertile ids; its last
//      CTA turns the histogram into the per-supertile list ranges and the slice plan of the fine
//      level.  One stable radix pass on the supertile id (two above 256 supertiles) turns the
//      instances into per-supertile lists in depth order.
//   3. FINE level: every supertile list is cut into slices of 64 entries, one warp per slice.
//      count  : per slice, how many entries cover each of the supertile's 64 tiles (lane = tile owner,
//               entries broadcast one by one in list order)
//      scan   : per (supertile, tile) exclusive prefix over the slices; the last CTA to finish scans
//               the per-tile totals in tile-id order, which gives every tile's start in `point_list`
//               and the tile `ranges` (reference identifyTileRanges, rasterizer_impl.cu:116-138:
//               empty tiles stay (0,0))
//      scatter: the count loop again; every tile owner appends the ids covering its tile at its running
//               position, so ids are written straight to their final place in stable order.
//
// Because (1) is stable in id, (2) is stable, and slices / lanes are walked in list order, instances
// inside a tile end up ordered by (depth bits, id): exactly the reference's sorted `point_list`.
// Traffic is ~60 P + 16 R1 + 4 R bytes instead of the reference's 152 R.
//
// NO HOST KNOWLEDGE OF COUNTS.  Every kernel here takes the CAPACITIES of its arrays from the host
// (P, R1_cap, R_cap) and reads the actual counts (V visible Gaussians, R1 supertile instances, the
// depth-key range) from the geometry header on the device.  Grids are sized by capacity; CTAs whose
// work unit lies beyond the actual count leave at once.  Writes are guarded by the capacities, so a
// forward whose capacities turn out too small is memory-safe; it raises the overflow word of the
// header and the caller runs it again with the exact sizes (api.cu).
//
// The radix sort is hand-written, one kernel per digit ("onesweep"): a histogram kernel counts all
// digits of all passes in one read of the keys; a pass then ranks a 4096-key tile (warp match_any +
// shared atomics), publishes the tile's digit counts, finds its global offsets by decoupled look-back
// over the predecessor tiles' published counts (tile ids are handed out by an atomic ticket, so a CTA
// only ever waits for CTAs that are already running) and scatters through shared memory so that global
// writes are coalesced per digit run.  Measured against cub::DeviceRadixSort::SortPairs on the same
// problem: tools/sort_vs_cub.cu, profiles/.
#include <algorithm>

#include "common.cuh"
#include "kernels.h"

namespace brs {

namespace {

constexpr int OS_THREADS = 256;
constexpr int OS_WARPS = OS_THREADS / 32;
constexpr int OS_ITEMS_MAX = 16; // keys per thread: 16 (4096-pair tiles) for big inputs, fewer for small ones, see sort_items()
constexpr int RADIX = 256;

constexpr uint32_t FLAG_LOCAL = 1u << 30; // tile-local count published
constexpr uint32_t FLAG_INCL = 2u << 30;  // inclusive prefix published
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
// s_warp must hold 8 uint32.  Contains two __syncthreads.
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
	for (int w = 0; w < OS_WARPS; w++) {
		const uint32_t c = s_warp[w];
		if ((uint32_t)w < warp)
			warp_off += c;
		tot += c;
	}
	__syncthreads();
	total = tot;
	return warp_off + incl - v;
}

// The depth sort's key transform lives on the device: bias = (min visible key) & ~255, so that the
// first digit (key bits 0..7) does not depend on it.  hdr[2] holds max(~key) = ~min.
__device__ __forceinline__ uint32_t depth_bias(const uint32_t* hdr)
{
	const uint32_t V = hdr[HDR_V];
	return V ? ((~hdr[HDR_KEY_INVMIN]) & ~0xFFu) : 0u;
}

// ---- histogram of every digit of every pass, one read of the keys -----------------------------------
struct HistArgs {
	const uint32_t* keys;
	uint32_t n;              // capacity
	const uint32_t* n_ptr;   // actual count on the device (nullptr: n)
	const uint32_t* hdr;     // geometry header (depth sort: bias + overflow bookkeeping) or nullptr
	uint32_t* hdr_out;       // same header, writable (overflow word), or nullptr
	int passes;
	int shift[4], bits[4];
	int drop;                // keys equal to DEPTH_KEY_CULLED do not take part
	uint32_t R_cap, R1_cap, V_cap; // for the overflow word (depth sort only)
	uint32_t* overflow_accum; // optional: the overflow bits are also OR-ed into this word
	uint32_t* hist;          // [passes][RADIX], pre-zeroed
};

__global__ void __launch_bounds__(OS_THREADS) radix_hist_kernel(HistArgs a)
{
	__shared__ uint32_t s_hist[4][RADIX];
	const uint32_t tid = threadIdx.x;
	for (int i = tid; i < 4 * RADIX; i += OS_THREADS)
		(&s_hist[0][0])[i] = 0;
	__syncthreads();
	const uint32_t n = a.n_ptr ? min(a.n, __ldg(a.n_ptr)) : a.n;
	const uint32_t bias = a.hdr ? depth_bias(a.hdr) : 0u;
	if (a.hdr_out != nullptr && blockIdx.x == 0 && tid == 0) {
		// overflow bookkeeping of the optimistic forward: do the capacities and the planned key bits hold?
		const uint32_t V = a.hdr[HDR_V];
		uint32_t nbits = 8;
		if (V) {
			const uint32_t span = a.hdr[HDR_KEY_MAX] - bias;
			while (nbits < 32 && (span >> nbits) != 0u)
				nbits++;
		}
		int planned = 0;
		for (int p = 0; p < a.passes; p++)
			planned = max(planned, a.shift[p] + a.bits[p]);
		uint32_t flags = 0;
		if (a.hdr[HDR_R] > a.R_cap)
			flags |= 1u;
		if (a.hdr[HDR_R1] > a.R1_cap)
			flags |= 2u;
		if ((int)nbits > planned)
			flags |= 4u;
		if (V > a.V_cap)
			flags |= 8u;
		a.hdr_out[HDR_OVERFLOW] = flags;
		a.hdr_out[HDR_KEY_BITS] = nbits;
		if (a.overflow_accum != nullptr && flags != 0u)
			atomicOr(a.overflow_accum, flags);
	}
	// four independent loads in flight per thread (a plain grid-stride loop exposes one load latency per key)
	const uint32_t stride = gridDim.x * OS_THREADS;
	for (uint32_t i0 = blockIdx.x * OS_THREADS + tid; i0 < n; i0 += 4 * stride) {
		uint32_t key[4];
#pragma unroll
		for (int j = 0; j < 4; j++) {
			const uint32_t i = i0 + j * stride;
			key[j] = (i >= i0 && i < n) ? __ldg(a.keys + i) : DEPTH_KEY_CULLED; // i >= i0: no 32-bit wrap-around
		}
#pragma unroll
		for (int j = 0; j < 4; j++) {
			const uint32_t i = i0 + j * stride;
			if (!(i >= i0 && i < n) || (a.drop && key[j] == DEPTH_KEY_CULLED))
				continue;
			const uint32_t k = key[j] - bias;
#pragma unroll
			for (int p = 0; p < 4; p++)
				if (p < a.passes)
					atomicAdd(&s_hist[p][(k >> a.shift[p]) & ((1u << a.bits[p]) - 1u)], 1u);
		}
	}
	__syncthreads();
	for (int i = tid; i < a.passes * RADIX; i += OS_THREADS) {
		const uint32_t c = (&s_hist[0][0])[i];
		if (c)
			atomicAdd(a.hist + i, c);
	}
}

inline uint32_t hist_blocks(size_t n) // 2048 keys per CTA, at most 4 CTAs per SM (measured: 296 ... 1184 CTAs within 3 %)
{
	return (uint32_t)std::max<size_t>(1, std::min<size_t>((n + 2047) / 2048, 592));
}

// ---- one radix pass ---------------------------------------------------------------------------------
struct PassArgs {
	const uint32_t* kin;
	const uint32_t* vin;     // nullptr: iota
	uint32_t* kout;          // nullptr: keys are not written (nobody reads them after the last pass)
	uint32_t* vout;
	uint32_t n;              // capacity
	const uint32_t* n_ptr;   // actual count (nullptr: n)
	uint32_t out_cap;        // entries of kout / vout: positions beyond it are dropped (only reachable after a capacity
	                         // overflow, when the digit totals count keys the pass does not see)
	const uint32_t* hdr;     // depth sort: bias from the header; nullptr: bias 0
	int shift, bits;
	int drop;
	const uint32_t* hist;    // [RADIX] digit totals of this pass, or (fold_n > 0) totals per full key value
	uint32_t fold_n;         // > 0: hist holds one count per key value 0..fold_n-1; fold it onto this pass's digit
	uint32_t* status;        // [tiles][RADIX], pre-zeroed
	uint32_t* ticket;        // pre-zeroed
};

template <int OS_ITEMS>
__global__ void __launch_bounds__(OS_THREADS) onesweep_kernel(PassArgs a)
{
	constexpr int OS_TILE = OS_THREADS * OS_ITEMS;
	__shared__ uint32_t s_cnt[OS_WARPS][RADIX];
	__shared__ uint32_t s_keys[OS_TILE];
	__shared__ uint32_t s_vals[OS_TILE];
	__shared__ uint32_t s_digit_start[RADIX];
	__shared__ uint32_t s_goff[RADIX];
	__shared__ uint32_t s_fold[RADIX];
	__shared__ uint32_t s_warp[OS_WARPS];
	__shared__ uint32_t s_tile;

	const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const uint32_t radix = 1u << a.bits, mask = radix - 1u;
	if (tid == 0)
		s_tile = atomicAdd(a.ticket, 1u);
	for (int i = tid; i < OS_WARPS * RADIX; i += OS_THREADS)
		(&s_cnt[0][0])[i] = 0;
	s_fold[tid] = 0;
	__syncthreads();
	const uint32_t n = a.n_ptr ? min(a.n, __ldg(a.n_ptr)) : a.n;
	const uint32_t tiles = (n + OS_TILE - 1) / OS_TILE;
	const uint32_t tile = s_tile;
	if (tile >= tiles)
		return;
	const uint32_t bias = a.hdr ? depth_bias(a.hdr) : 0u;
	const uint32_t base = tile * OS_TILE;
	const uint32_t valid_count = min((uint32_t)OS_TILE, n - base);

	// load keys (warp-striped: coalesced, and rank order == index order)
	// (values are fetched here as well, unconditionally: fetched where they are stored, behind the `valid`
	// branch, the OS_ITEMS loads would be issued one after the other, each exposing its full latency)
	uint32_t key[OS_ITEMS], val[OS_ITEMS], rank[OS_ITEMS];
	const uint32_t wbase = warp * (32 * OS_ITEMS) + lane;
#pragma unroll
	for (int i = 0; i < OS_ITEMS; i++) {
		const uint32_t li = wbase + i * 32;
		key[i] = (li < valid_count) ? __ldg(a.kin + base + li) : 0xffffffffu;
	}
#pragma unroll
	for (int i = 0; i < OS_ITEMS; i++) {
		const uint32_t li = wbase + i * 32;
		val[i] = base + li;
		if (a.vin != nullptr && li < valid_count)
			val[i] = __ldg(a.vin + base + li);
	}
	// digit totals over the whole input (independent of the ranking below)
	uint32_t digit_total = 0;
	if (a.fold_n) {
		for (uint32_t s = tid; s < a.fold_n; s += OS_THREADS) {
			const uint32_t c = ld_relaxed(a.hist + s);
			if (c)
				atomicAdd(&s_fold[(s >> a.shift) & mask], c);
		}
	} else if (tid < radix) {
		digit_total = ld_relaxed(a.hist + tid);
	}

	// rank inside the warp (stable).  match_any groups equal digits; the group's first lane bumps the
	// warp-private counter with a shared-memory atomic that returns the old value.  Same-address
	// atomics of one warp complete in program order, so no barrier is needed and all OS_ITEMS
	// chains (MATCH -> ATOMS -> SHFL) overlap.
	uint32_t* my_cnt = s_cnt[warp];
#pragma unroll
	for (int i = 0; i < OS_ITEMS; i++) {
		const bool valid = (wbase + i * 32) < valid_count && !(a.drop && key[i] == DEPTH_KEY_CULLED);
		const uint32_t d = ((key[i] - bias) >> a.shift) & mask;
		const uint32_t peers = __match_any_sync(0xffffffffu, valid ? d : 0xffffffffu);
		const int leader = __ffs(peers) - 1;
		uint32_t old = 0;
		if (valid && (int)lane == leader)
			old = atomicAdd(my_cnt + d, (uint32_t)__popc(peers));
		old = __shfl_sync(0xffffffffu, old, leader);
		rank[i] = old + __popc(peers & lanemask_lt());
	}
	__syncthreads();
	if (a.fold_n && tid < radix)
		digit_total = s_fold[tid];

	// per digit: exclusive offsets of the warps, tile total; publish the tile's digit counts at once so
	// that successors can start summing them
	uint32_t total = 0;
	if (tid < radix) {
#pragma unroll
		for (int w = 0; w < OS_WARPS; w++) {
			const uint32_t c = s_cnt[w][tid];
			s_cnt[w][tid] = total;
			total += c;
		}
		st_relaxed(a.status + (size_t)tile * RADIX + tid, (tile == 0 ? FLAG_INCL : FLAG_LOCAL) | total);
	}
	uint32_t dummy, kept; // kept = keys of this tile that take part (all valid ones unless `drop`)
	const uint32_t digit_start = block_exclusive_scan_256(total, s_warp, kept);
	const uint32_t bin_base = block_exclusive_scan_256(digit_total, s_warp, dummy);
	if (tid < radix)
		s_digit_start[tid] = digit_start;
	__syncthreads();

	// scatter into shared memory in sorted-by-digit order; this also gives the predecessors time to publish
#pragma unroll
	for (int i = 0; i < OS_ITEMS; i++) {
		const uint32_t li = wbase + i * 32;
		if (li < valid_count && !(a.drop && key[i] == DEPTH_KEY_CULLED)) {
			const uint32_t d = ((key[i] - bias) >> a.shift) & mask;
			const uint32_t pos = s_digit_start[d] + s_cnt[warp][d] + rank[i];
			s_keys[pos] = key[i];
			s_vals[pos] = val[i];
		}
	}

	// decoupled look-back: thread d sums digit d over the predecessor tiles until it meets an inclusive prefix
	if (tid < radix) {
		uint32_t excl = 0;
		int p = (int)tile - 1;
		while (p >= 0) {
			const uint32_t v = ld_relaxed(a.status + (size_t)p * RADIX + tid);
			const uint32_t f = v & FLAG_MASK;
			if (f == 0)
				continue; // not published yet (that tile holds an earlier ticket: it is running)
			excl += v & VALUE_MASK;
			if (f == FLAG_INCL)
				break;
			p--;
		}
		if (tile > 0)
			st_relaxed(a.status + (size_t)tile * RADIX + tid, FLAG_INCL | (excl + total));
		s_goff[tid] = bin_base + excl - digit_start; // global position = s_goff[d] + position in tile
	}
	__syncthreads();

	// coalesced write-out: consecutive threads hold consecutive positions of a digit run
#pragma unroll
	for (int i = 0; i < OS_ITEMS; i++) {
		const uint32_t j = tid + i * OS_THREADS;
		if (j < kept) {
			const uint32_t k = s_keys[j];
			const uint32_t g = s_goff[((k - bias) >> a.shift) & mask] + j;
			if (g < a.out_cap) {
				if (a.kout)
					a.kout[g] = k;
				a.vout[g] = s_vals[j];
			}
		}
	}
}

// Keys per thread of a pass over (at most) n pairs.  A tile's work is a latency chain (ticket -> loads -> ranking ->
// look-back -> scatter); with 4096-pair tiles an input of 100 K pairs is 25 CTAs on 148 SMs, each walking 16 keys per
// thread through that chain.  Smaller tiles spread a small input over the SMs and shorten the chain; big inputs keep
// the big tiles (fewer status words, fewer look-back steps per key).  Chosen from the CAPACITY, which the host knows
// when it sizes the status arrays.
inline int sort_items(size_t n)
{
	static const int forced = [] { const char* e = getenv("BRS_SORT_ITEMS"); return e ? atoi(e) : 0; }();
	if (forced == 4 || forced == 8 || forced == 16)
		return forced;
	return n <= 128u * 1024u ? 4 : (n <= 384u * 1024u ? 8 : 16); // measured against CUB at 50 K ... 3 M pairs
}
inline uint32_t sort_tiles(size_t n)
{
	const size_t tile = (size_t)OS_THREADS * sort_items(n);
	return (uint32_t)((n + tile - 1) / tile);
}
inline void launch_onesweep(const PassArgs& a, size_t capacity, cudaStream_t stream)
{
	const uint32_t tiles = sort_tiles(capacity);
	if (tiles == 0)
		return; // capacity 0 (nothing visible, known exactly): nothing to sort
	count_launch();
	switch (sort_items(capacity)) {
	case 4: onesweep_kernel<4><<<tiles, OS_THREADS, 0, stream>>>(a); break;
	case 8: onesweep_kernel<8><<<tiles, OS_THREADS, 0, stream>>>(a); break;
	default: onesweep_kernel<16><<<tiles, OS_THREADS, 0, stream>>>(a); break;
	}
}

// ---- fused scan + emission --------------------------------------------------------------------------

constexpr int EMIT_THREADS = 256;
constexpr int EMIT_ITEMS = 4;
constexpr int EMIT_TILE = EMIT_THREADS * EMIT_ITEMS; // Gaussians per block
constexpr int CELL_HIST_MAX = 2048;                  // supertile ids histogrammed in shared memory

// Tile rectangle -> rectangle in cells of (1 << shift) tiles; an empty rectangle stays empty.
__device__ __forceinline__ uint2 coarsen_rect(uint2 r, uint32_t shift)
{
	const uint32_t x0 = r.x & 0xffffu, x1 = r.x >> 16, y0 = r.y & 0xffffu, y1 = r.y >> 16;
	if (x1 <= x0 || y1 <= y0)
		return make_uint2(0u, 0u);
	return make_uint2((x0 >> shift) | ((((x1 - 1) >> shift) + 1) << 16), (y0 >> shift) | ((((y1 - 1) >> shift) + 1) << 16));
}

struct EmitArgs {
	const uint32_t* order;   // ids in depth order
	const uint2* rect;
	const uint32_t* hdr;     // V = min(hdr[HDR_V], V_cap)
	uint32_t V_cap;          // entries of `order` the depth sort can have written
	uint32_t P;              // ids are < P
	uint32_t shift, ns_x, ns;
	uint32_t* cell_keys;     // [R1_cap]
	uint32_t* cell_ids;      // [R1_cap]
	uint32_t R1_cap;
	uint32_t* status;        // [blocks], pre-zeroed
	uint32_t* ticket;        // pre-zeroed
	uint32_t* done;          // pre-zeroed
	uint32_t* cell_count;    // [ns], pre-zeroed
	uint2* coarse_ranges;    // [ns]      (written by the last CTA)
	uint32_t* slice_base;    // [ns + 1]  (written by the last CTA)
	uint32_t* n_instances;   // min(R1, R1_cap) (written by the last CTA)
};

// The last CTA of the emission: per-supertile totals -> list ranges in the sorted instance array and the
// fine level's slice plan (ceil(n_s / FINE_SLICE) slices per supertile, exclusive scan).
__device__ void emit_finalize(const EmitArgs& a, uint32_t* s_warp)
{
	uint32_t carry_inst = 0, carry_slices = 0;
	for (uint32_t s0 = 0; s0 < a.ns; s0 += EMIT_THREADS) {
		const uint32_t s = s0 + threadIdx.x;
		const uint32_t v = (s < a.ns) ? ld_relaxed(a.cell_count + s) : 0u;
		uint32_t tot_i, tot_s;
		const uint32_t ex_i = block_exclusive_scan_256(v, s_warp, tot_i);
		const uint32_t ex_s = block_exclusive_scan_256((v + FINE_SLICE - 1) / FINE_SLICE, s_warp, tot_s);
		if (s < a.ns) {
			a.coarse_ranges[s] = make_uint2(carry_inst + ex_i, carry_inst + ex_i + v);
			a.slice_base[s] = carry_slices + ex_s;
		}
		carry_inst += tot_i;
		carry_slices += tot_s;
	}
	if (threadIdx.x == 0) {
		a.slice_base[a.ns] = carry_slices;
		*a.n_instances = carry_inst;
	}
}

__global__ void __launch_bounds__(EMIT_THREADS) emit_kernel(EmitArgs a)
{
	__shared__ uint32_t s_incl[EMIT_TILE]; // inclusive instance offsets inside this block
	__shared__ uint32_t s_id[EMIT_TILE];
	__shared__ uint2 s_rect[EMIT_TILE];
	__shared__ uint32_t s_hist[CELL_HIST_MAX];
	__shared__ uint32_t s_warp[OS_WARPS];
	__shared__ uint32_t s_bcast[2];

	const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	if (tid == 0)
		s_bcast[0] = atomicAdd(a.ticket, 1u);
	const bool smem_hist = a.ns <= CELL_HIST_MAX;
	if (smem_hist)
		for (uint32_t i = tid; i < a.ns; i += EMIT_THREADS)
			s_hist[i] = 0;
	__syncthreads();
	const uint32_t blk = s_bcast[0];
	const uint32_t V = min(__ldg(a.hdr + HDR_V), a.V_cap); // beyond the capacity: the overflow word is raised, the caller re-runs
	const uint32_t blocks = (V + EMIT_TILE - 1) / EMIT_TILE;
	if (blk >= blocks) {
		if (blocks == 0 && blk == 0)
			emit_finalize(a, s_warp); // nothing visible: all lists are empty
		return;
	}
	const uint32_t first = blk * EMIT_TILE + tid * EMIT_ITEMS;

	uint32_t cnt[EMIT_ITEMS];
	uint32_t tsum = 0;
#pragma unroll
	for (int k = 0; k < EMIT_ITEMS; k++) {
		const uint32_t s = first + k;
		uint32_t id = 0;
		uint2 r = make_uint2(0u, 0u);
		if (s < V) {
			id = __ldg(a.order + s);
			// after a V_cap overflow the sort leaves unwritten slots in `order`: any value may sit there
			if (id < a.P)
				r = coarsen_rect(__ldg(a.rect + id), a.shift);
			else
				id = 0;
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
			st_relaxed(a.status + blk, (blk == 0 ? FLAG_INCL : FLAG_LOCAL) | block_total);
		uint32_t excl = 0;
		int p = (int)blk - 1;
		while (p >= 0) {
			const int q = p - (int)lane;
			const uint32_t v = (q >= 0) ? ld_relaxed(a.status + q) : FLAG_INCL;
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
				st_relaxed(a.status + blk, FLAG_INCL | (excl + block_total));
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
		if (g < a.R1_cap) { // beyond the capacity: dropped (the header's overflow word is raised, the caller re-runs)
			const uint32_t cell = (y0 + row) * a.ns_x + x0 + (k - row * w);
			a.cell_keys[g] = cell;
			a.cell_ids[g] = s_id[lo];
			if (smem_hist)
				atomicAdd(&s_hist[cell], 1u);
			else
				atomicAdd(a.cell_count + cell, 1u);
		}
	}
	__syncthreads();
	if (smem_hist)
		for (uint32_t i = tid; i < a.ns; i += EMIT_THREADS) {
			const uint32_t c = s_hist[i];
			if (c)
				atomicAdd(a.cell_count + i, c);
		}
	// last CTA done -> finalize
	__threadfence();
	__syncthreads();
	if (tid == 0)
		s_bcast[0] = atomicAdd(a.done, 1u);
	__syncthreads();
	if (s_bcast[0] == blocks - 1) {
		__threadfence();
		emit_finalize(a, s_warp);
	}
}

// ---- fine level: supertile lists -> per-tile lists --------------------------------------------------

constexpr int FINE_WARPS = 8; // warps (= slices) per CTA of the count / scatter kernels

__device__ __forceinline__ uint32_t spread4(uint32_t b) // bit i of b (i < 4) -> bit 8 i
{
	return (b * 0x00204081u) & 0x01010101u;
}

// In-warp transpose of a 32 x 32 bit matrix: lane i holds row i (bit j = element (i, j)); afterwards lane i
// holds column i (bit j = element (j, i)).  Five butterfly stages, each swapping the off-diagonal k x k blocks.
__device__ __forceinline__ uint32_t transpose32(uint32_t v, uint32_t lane)
{
#pragma unroll
	for (int s = 0; s < 5; s++) {
		const uint32_t k = 16u >> s;
		const uint32_t m = s == 0 ? 0x0000FFFFu : (s == 1 ? 0x00FF00FFu : (s == 2 ? 0x0F0F0F0Fu : (s == 3 ? 0x33333333u : 0x55555555u)));
		const uint32_t x = __shfl_xor_sync(0xffffffffu, v, k);
		v = (lane & k) ? ((v & ~m) | ((x >> k) & m)) : ((v & m) | ((x << k) & ~m));
	}
	return v;
}

// One warp per slice of <= FINE_SLICE consecutive entries of one supertile's list.  Entries are read 32
// at a time (lane = entry, list order): each lane builds the 64-bit mask of the supertile's tiles its
// rectangle covers.  Two in-warp bit-matrix transposes flip the roles: lane t now OWNS local tiles t (rows
// 0..3) and t + 32 (rows 4..7) and holds, per tile, the 32-bit set of entries that cover it, in list order.
// COUNT pass: a popcount.  SCATTER pass: the owner walks its set bits in ascending order and appends the
// ids (fetched from the entry's lane by shuffle) at its running position - walking the bits in order IS the
// stable order.
template <bool SCATTER>
__global__ void __launch_bounds__(FINE_WARPS * 32)
    fine_kernel(const uint32_t* __restrict__ coarse_list, const uint2* __restrict__ coarse_ranges,
                const uint32_t* __restrict__ slice_base, const uint2* __restrict__ rect, uint32_t ns, uint32_t ns_x,
                uint32_t grid_x, uint32_t grid_y, uint32_t* __restrict__ table, const uint32_t* __restrict__ tile_start,
                uint32_t* __restrict__ point_list, uint32_t R_cap)
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

	// both 32-entry chunks of the slice are fetched up front (id -> rectangle is a dependent gather)
	constexpr int CHUNKS = FINE_SLICE / 32;
	uint32_t ids[CHUNKS], los[CHUNKS], his[CHUNKS];
#pragma unroll
	for (int c = 0; c < CHUNKS; c++) {
		const uint32_t e = begin + 32 * c + lane;
		ids[c] = (e < end) ? __ldg(coarse_list + e) : 0u;
	}
#pragma unroll
	for (int c = 0; c < CHUNKS; c++) {
		const uint32_t e = begin + 32 * c + lane;
		los[c] = 0;
		his[c] = 0;
		if (e < end) {
			const uint2 r = __ldg(rect + ids[c]);
			const int x0 = (int)(r.x & 0xffffu) - (int)tx0, x1 = (int)(r.x >> 16) - (int)tx0;
			const int y0 = (int)(r.y & 0xffffu) - (int)ty0, y1 = (int)(r.y >> 16) - (int)ty0;
			const uint32_t cx0 = (uint32_t)max(x0, 0), cx1 = (uint32_t)min(x1, 8);
			const uint32_t cy0 = (uint32_t)max(y0, 0), cy1 = (uint32_t)min(y1, 8);
			if (cx1 > cx0 && cy1 > cy0) {
				const uint32_t cols = ((1u << cx1) - (1u << cx0)) & 0xffu;
				const uint32_t rows = ((1u << cy1) - (1u << cy0)) & 0xffu;
				los[c] = cols * spread4(rows & 15u);
				his[c] = cols * spread4(rows >> 4);
			}
		}
	}
#pragma unroll
	for (int c = 0; c < CHUNKS; c++) {
		if (begin + 32 * c >= end)
			break;
		// bit j of m0 / m1: entry j of this chunk covers this lane's tile (rows 0..3 / rows 4..7)
		uint32_t m0 = transpose32(los[c], lane), m1 = transpose32(his[c], lane);
		if (SCATTER) {
			const uint32_t most = __reduce_max_sync(0xffffffffu, max(__popc(m0), __popc(m1)));
			for (uint32_t it = 0; it < most; it++) {
				const uint32_t j0 = m0 ? (uint32_t)__ffs(m0) - 1u : 0u, j1 = m1 ? (uint32_t)__ffs(m1) - 1u : 0u;
				const uint32_t id0 = __shfl_sync(0xffffffffu, ids[c], j0);
				const uint32_t id1 = __shfl_sync(0xffffffffu, ids[c], j1);
				if (m0) {
					if (run0 < R_cap)
						point_list[run0] = id0;
					run0++;
					m0 &= m0 - 1u;
				}
				if (m1) {
					if (run1 < R_cap)
						point_list[run1] = id1;
					run1++;
					m1 &= m1 - 1u;
				}
			}
		} else {
			run0 += __popc(m0);
			run1 += __popc(m1);
		}
	}
	if (!SCATTER) {
		table[(size_t)sl * ST_TILES + lane] = run0;
		table[(size_t)sl * ST_TILES + 32 + lane] = run1;
	}
}

// Per supertile (one CTA, FS_GROUPS groups x 64 tiles): exclusive prefix of the slice counts along the
// slices, in place, and the per-tile totals into tile_count[tile id].  The last CTA to finish then
// scans the per-tile totals in tile-id order -> tile_start, and the tile ranges (reference
// identifyTileRanges: tiles without instances keep (0, 0)); ranges are clamped to the capacity of
// `point_list` so that an overflowing forward stays memory-safe.
constexpr int FS_GROUPS = 16; // a supertile's slices are cut into this many runs, one per 64-thread group
__global__ void __launch_bounds__(FS_GROUPS * ST_TILES)
    fine_scan_kernel(uint32_t* __restrict__ table, const uint32_t* __restrict__ slice_base, uint32_t ns_x, uint32_t grid_x,
                     uint32_t grid_y, uint32_t* __restrict__ tile_count, uint32_t* __restrict__ tile_start,
                     uint2* __restrict__ ranges, uint32_t R_cap, uint32_t* done)
{
	__shared__ uint32_t s_part[FS_GROUPS][ST_TILES];
	__shared__ uint32_t s_warp[OS_WARPS];
	__shared__ uint32_t s_last;
	const uint32_t st = blockIdx.x;
	const uint32_t first = __ldg(slice_base + st), nsl = __ldg(slice_base + st + 1) - first;
	const uint32_t t = threadIdx.x & 63, q = threadIdx.x >> 6;
	const uint32_t per = (nsl + FS_GROUPS - 1) / FS_GROUPS;
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
	for (uint32_t k = 0; k < FS_GROUPS; k++) {
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
			st_relaxed(tile_count + y * grid_x + x, total);
	}
	__threadfence();
	__syncthreads();
	if (threadIdx.x == 0)
		s_last = atomicAdd(done, 1u);
	__syncthreads();
	if (s_last != gridDim.x - 1 || threadIdx.x >= 256)
		return;
	__threadfence();
	// exclusive scan over all tiles in tile-id order by the CTA's first 256 threads (named barrier 1: the other
	// threads have left): every thread takes a contiguous run of tiles
	const uint32_t num_tiles = grid_x * grid_y;
	const uint32_t chunk = (num_tiles + 255) / 256;
	const uint32_t t0 = min(threadIdx.x * chunk, num_tiles), t1 = min(t0 + chunk, num_tiles);
	// (plain L2 loads after the fence: unlike volatile ones they are issued back to back)
	uint32_t local = 0;
#pragma unroll 8
	for (uint32_t i = t0; i < t1; i++)
		local += __ldcg(tile_count + i);
	uint32_t grand;
	uint32_t start = block_exclusive_scan_256(local, s_warp, grand);
#pragma unroll 8
	for (uint32_t i = t0; i < t1; i++) {
		const uint32_t v = __ldcg(tile_count + i);
		tile_start[i] = start;
		ranges[i] = v ? make_uint2(min(start, R_cap), min(start + v, R_cap)) : make_uint2(0u, 0u);
		start += v;
	}
}

template <class T>
T* carve(char*& p, size_t count)
{
	T* r = reinterpret_cast<T*>(p);
	p += align_up(sizeof(T) * (count ? count : 1), 256);
	return r;
}

} // namespace

// ---- host side --------------------------------------------------------------------------------------

namespace {
// Both carvers walk the zeroed and the plain area with the same code that sizes them (nullptr base).
struct DepthCarve {
	DepthScratch s;
	size_t zero_bytes, plain_bytes;
};
DepthCarve carve_depth(void* zeroed, void* plain, size_t P, size_t V_cap, int passes)
{
	DepthCarve c{};
	// the first pass works on the P keys, the later ones on the (at most V_cap) visible ones it kept
	const size_t tiles = sort_tiles(P), tiles_rest = sort_tiles(V_cap < P ? V_cap : P);
	char* z = static_cast<char*>(zeroed);
	c.s.ctl = carve<uint32_t>(z, DCTL_WORDS);
	c.s.hist = carve<uint32_t>(z, 4 * RADIX);
	c.s.status = carve<uint32_t>(z, (tiles + (size_t)(passes > 1 ? passes - 1 : 0) * tiles_rest) * RADIX);
	c.s.tiles_rest = (uint32_t)tiles_rest;
	c.zero_bytes = (size_t)(z - static_cast<char*>(zeroed));
	char* p = static_cast<char*>(plain);
	for (int i = 0; i < 2; i++) {
		c.s.keys[i] = carve<uint32_t>(p, P);
		c.s.vals[i] = carve<uint32_t>(p, P);
	}
	c.plain_bytes = (size_t)(p - static_cast<char*>(plain));
	c.s.tiles = (uint32_t)tiles;
	return c;
}
struct InstCarve {
	InstScratch s;
	size_t zero_bytes, plain_bytes;
};
InstCarve carve_inst(void* zeroed, void* plain, size_t P, size_t R1_cap, uint32_t grid_x, uint32_t grid_y)
{
	InstCarve c{};
	const size_t ns = (size_t)supertiles(grid_x) * supertiles(grid_y);
	const size_t num_tiles = (size_t)grid_x * grid_y;
	const size_t tiles_r = sort_tiles(R1_cap);
	const size_t emit_blocks = (P + EMIT_TILE - 1) / EMIT_TILE;
	const size_t max_slices = R1_cap / FINE_SLICE + ns;
	const int coarse_passes = ns > RADIX ? 2 : 1;
	char* z = static_cast<char*>(zeroed);
	c.s.ctl = carve<uint32_t>(z, ICTL_WORDS);
	c.s.cell_count = carve<uint32_t>(z, ns);
	c.s.emit_status = carve<uint32_t>(z, emit_blocks);
	c.s.coarse_status = carve<uint32_t>(z, (size_t)coarse_passes * tiles_r * RADIX);
	c.zero_bytes = (size_t)(z - static_cast<char*>(zeroed));
	char* p = static_cast<char*>(plain);
	c.s.cell_keys = carve<uint32_t>(p, R1_cap);
	c.s.cell_ids = carve<uint32_t>(p, R1_cap);
	c.s.tmp_keys = carve<uint32_t>(p, coarse_passes > 1 ? R1_cap : 0);
	c.s.tmp_ids = carve<uint32_t>(p, coarse_passes > 1 ? R1_cap : 0);
	c.s.coarse_list = carve<uint32_t>(p, R1_cap);
	c.s.coarse_ranges = carve<uint2>(p, ns);
	c.s.slice_base = carve<uint32_t>(p, ns + 1);
	c.s.tile_count = carve<uint32_t>(p, num_tiles);
	c.s.tile_start = carve<uint32_t>(p, num_tiles);
	c.s.table = carve<uint32_t>(p, ST_TILES * max_slices);
	c.plain_bytes = (size_t)(p - static_cast<char*>(plain));
	c.s.coarse_tiles = (uint32_t)tiles_r;
	c.s.coarse_passes = coarse_passes;
	return c;
}
} // namespace

size_t depth_zero_bytes(size_t P, size_t V_cap, int passes) { return carve_depth(nullptr, nullptr, P, V_cap, passes).zero_bytes; }
size_t depth_plain_bytes(size_t P) { return carve_depth(nullptr, nullptr, P, P, 1).plain_bytes; }
DepthScratch carve_depth_scratch(void* zeroed, void* plain, size_t P, size_t V_cap, int passes)
{
	return carve_depth(zeroed, plain, P, V_cap, passes).s;
}
size_t inst_zero_bytes(size_t P, size_t R1_cap, uint32_t grid_x, uint32_t grid_y)
{
	return carve_inst(nullptr, nullptr, P, R1_cap, grid_x, grid_y).zero_bytes;
}
size_t inst_plain_bytes(size_t P, size_t R1_cap, uint32_t grid_x, uint32_t grid_y)
{
	return carve_inst(nullptr, nullptr, P, R1_cap, grid_x, grid_y).plain_bytes;
}
InstScratch carve_inst_scratch(void* zeroed, void* plain, size_t P, size_t R1_cap, uint32_t grid_x, uint32_t grid_y)
{
	return carve_inst(zeroed, plain, P, R1_cap, grid_x, grid_y).s;
}

namespace {
PassArgs depth_pass_args(const BinPlan& pl, int p)
{
	const DepthScratch& d = pl.d;
	PassArgs a{};
	const bool last = p == pl.depth_passes - 1;
	a.kin = p == 0 ? pl.depth_key : d.keys[(p - 1) & 1];
	a.vin = p == 0 ? nullptr : d.vals[(p - 1) & 1];
	a.kout = last ? nullptr : d.keys[p & 1];
	a.vout = last ? pl.order : d.vals[p & 1];
	a.n = p == 0 ? pl.P : min(pl.P, pl.V_cap);
	a.out_cap = pl.P;
	a.n_ptr = p == 0 ? nullptr : pl.hdr + HDR_V; // the first pass drops the culled Gaussians
	a.hdr = pl.hdr;
	a.shift = 8 * p;
	a.bits = 8;
	a.drop = p == 0;
	a.hist = d.hist + p * RADIX;
	a.fold_n = 0;
	a.status = d.status + (p == 0 ? 0 : ((size_t)d.tiles + (size_t)(p - 1) * d.tiles_rest) * RADIX);
	a.ticket = d.ctl + DCTL_TICKET + p;
	return a;
}
} // namespace

// Depth sort of the visible Gaussians: (depth_key - bias, id) on `depth_passes` 8-bit digits -> `order`.
// begin: histogram of the digits of `planned_passes` passes (+ the overflow word of the header) and the
// first pass, none of which needs the host to know anything; rest: passes 1 .. depth_passes - 1.
// With depth_passes == 1 the first pass already writes `order`, so depth_passes must be final by then
// (pl.depth_passes == 1 is only planned from a previous call's key range; the exact path plans >= 2).
cudaError_t launch_depth_sort_begin(const BinPlan& pl, int planned_passes, cudaStream_t stream)
{
	if (pl.P == 0)
		return cudaSuccess;
	HistArgs h{};
	h.keys = pl.depth_key;
	h.n = pl.P;
	h.n_ptr = nullptr;
	h.hdr = pl.hdr;
	h.hdr_out = pl.hdr;
	h.passes = planned_passes;
	for (int p = 0; p < 4; p++) {
		h.shift[p] = 8 * p;
		h.bits[p] = 8;
	}
	h.drop = 1;
	h.R_cap = pl.R_cap;
	h.R1_cap = pl.R1_cap;
	h.V_cap = pl.V_cap;
	h.overflow_accum = pl.overflow_accum;
	h.hist = pl.d.hist;
	const uint32_t hist_grid = hist_blocks(pl.P);
	radix_hist_kernel<<<hist_grid, OS_THREADS, 0, stream>>>(h);
	count_launch();
	launch_onesweep(depth_pass_args(pl, 0), pl.P, stream);
	return cudaGetLastError();
}

cudaError_t launch_depth_sort_rest(const BinPlan& pl, cudaStream_t stream)
{
	if (pl.P == 0)
		return cudaSuccess;
	for (int p = 1; p < pl.depth_passes; p++) {
		launch_onesweep(depth_pass_args(pl, p), min(pl.P, pl.V_cap), stream);
	}
	return cudaGetLastError();
}

cudaError_t launch_emit(const BinPlan& pl, cudaStream_t stream)
{
	const InstScratch& b = pl.i;
	if (pl.P == 0)
		return cudaSuccess;
	EmitArgs a{};
	a.order = pl.order;
	a.rect = pl.rect;
	a.hdr = pl.hdr;
	a.V_cap = min(pl.P, pl.V_cap);
	a.P = pl.P;
	a.shift = ST_SHIFT;
	a.ns_x = pl.ns_x;
	a.ns = pl.ns;
	a.cell_keys = b.cell_keys;
	a.cell_ids = b.cell_ids;
	a.R1_cap = pl.R1_cap;
	a.status = b.emit_status;
	a.ticket = b.ctl + ICTL_EMIT_TICKET;
	a.done = b.ctl + ICTL_EMIT_DONE;
	a.cell_count = b.cell_count;
	a.coarse_ranges = b.coarse_ranges;
	a.slice_base = b.slice_base;
	a.n_instances = b.ctl + ICTL_N_INSTANCES;
	const uint32_t blocks = max(1u, (uint32_t)((min(pl.P, pl.V_cap) + EMIT_TILE - 1) / EMIT_TILE)); // >= 1: an empty emission still writes the (empty) plan
	emit_kernel<<<blocks, EMIT_THREADS, 0, stream>>>(a);
	count_launch();
	return cudaGetLastError();
}

// Stable radix pass(es) on the supertile id: (cell_keys, cell_ids) -> coarse_list.
cudaError_t launch_coarse_sort(const BinPlan& pl, cudaStream_t stream)
{
	const InstScratch& b = pl.i;
	if (pl.P == 0 || pl.R1_cap == 0)
		return cudaSuccess;
	int nbits = 1;
	while ((1u << nbits) < pl.ns)
		nbits++;
	const int passes = b.coarse_passes;
	const int lo_bits = passes == 1 ? nbits : nbits / 2;
	if (nbits - lo_bits > 8 || lo_bits > 8)
		return cudaErrorInvalidValue; // more than 65536 supertiles (image beyond 32768 x 32768 px)
	for (int p = 0; p < passes; p++) {
		PassArgs a{};
		const bool last = p == passes - 1;
		a.kin = p == 0 ? b.cell_keys : b.tmp_keys;
		a.vin = p == 0 ? b.cell_ids : b.tmp_ids;
		a.kout = last ? nullptr : b.tmp_keys;
		a.vout = last ? b.coarse_list : b.tmp_ids;
		a.n = pl.R1_cap;
		a.out_cap = pl.R1_cap;
		a.n_ptr = b.ctl + ICTL_N_INSTANCES;
		a.hdr = nullptr;
		a.shift = p == 0 ? 0 : lo_bits;
		a.bits = p == 0 ? lo_bits : nbits - lo_bits;
		a.drop = 0;
		a.hist = b.cell_count;
		a.fold_n = pl.ns;
		a.status = b.coarse_status + (size_t)p * b.coarse_tiles * RADIX;
		a.ticket = b.ctl + ICTL_COARSE_TICKET + p;
		launch_onesweep(a, pl.R1_cap, stream);
	}
	return cudaGetLastError();
}

cudaError_t launch_fine_binning(const BinPlan& pl, cudaStream_t stream)
{
	const InstScratch& b = pl.i;
	const uint32_t num_tiles = pl.grid_x * pl.grid_y;
	if (num_tiles == 0 || pl.P == 0)
		return cudaSuccess;
	const uint32_t max_slices = (uint32_t)(pl.R1_cap / FINE_SLICE + pl.ns);
	const uint32_t blocks = (max_slices + FINE_WARPS - 1) / FINE_WARPS;
	if (pl.R1_cap > 0) {
		fine_kernel<false><<<blocks, FINE_WARPS * 32, 0, stream>>>(b.coarse_list, b.coarse_ranges, b.slice_base, pl.rect, pl.ns,
		                                                            pl.ns_x, pl.grid_x, pl.grid_y, b.table, nullptr, nullptr, 0u);
		count_launch();
	}
	fine_scan_kernel<<<pl.ns, FS_GROUPS * ST_TILES, 0, stream>>>(b.table, b.slice_base, pl.ns_x, pl.grid_x, pl.grid_y, b.tile_count,
	                                            b.tile_start, pl.ranges, pl.R_cap, b.ctl + ICTL_FINE_DONE);
	count_launch();
	if (pl.R1_cap > 0 && pl.R_cap > 0) {
		fine_kernel<true><<<blocks, FINE_WARPS * 32, 0, stream>>>(b.coarse_list, b.coarse_ranges, b.slice_base, pl.rect, pl.ns,
		                                                           pl.ns_x, pl.grid_x, pl.grid_y, b.table, b.tile_start,
		                                                           pl.point_list, pl.R_cap);
		count_launch();
	}
	return cudaGetLastError();
}

// ---- stand-alone stable sort of (u32, u32) pairs (C-ABI brs_sort_pairs_u32, tests, tools/sort_vs_cub) ----

namespace {
struct PairSortScratch {
	uint32_t* ctl;    // [64] tickets
	uint32_t* hist;   // [4][RADIX]
	uint32_t* status; // [4][tiles][RADIX]
	size_t zero_bytes;
	uint32_t* tmp_keys;
	uint32_t* tmp_vals;
	size_t total_bytes;
};
PairSortScratch carve_pair_sort(void* scratch, size_t n)
{
	PairSortScratch s{};
	char* p = static_cast<char*>(scratch);
	s.ctl = carve<uint32_t>(p, 64);
	s.hist = carve<uint32_t>(p, 4 * RADIX);
	s.status = carve<uint32_t>(p, 4 * (size_t)sort_tiles(n) * RADIX);
	s.zero_bytes = (size_t)(p - static_cast<char*>(scratch));
	s.tmp_keys = carve<uint32_t>(p, n);
	s.tmp_vals = carve<uint32_t>(p, n);
	s.total_bytes = (size_t)(p - static_cast<char*>(scratch));
	return s;
}
} // namespace

size_t sort_scratch_bytes(size_t n) { return carve_pair_sort(nullptr, n).total_bytes; }

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
	PairSortScratch s = carve_pair_sort(scratch, n);
	cudaError_t e = cudaMemsetAsync(scratch, 0, s.zero_bytes, stream);
	if (e != cudaSuccess)
		return e;
	HistArgs h{};
	h.keys = keys_in;
	h.n = (uint32_t)n;
	h.passes = passes;
	int shift = begin_bit;
	for (int p = 0; p < passes; p++) {
		h.shift[p] = shift;
		h.bits[p] = base_bits + (p < extra ? 1 : 0);
		shift += h.bits[p];
	}
	h.hist = s.hist;
	const uint32_t tiles = sort_tiles(n);
	radix_hist_kernel<<<hist_blocks(n), OS_THREADS, 0, stream>>>(h);
	count_launch();
	const uint32_t* kin = keys_in;
	const uint32_t* vin = vals_in;
	for (int p = 0; p < passes; p++) {
		const bool to_out = ((passes - 1 - p) & 1) == 0;
		PassArgs a{};
		a.kin = kin;
		a.vin = vin;
		a.kout = to_out ? keys_out : s.tmp_keys;
		a.vout = to_out ? vals_out : s.tmp_vals;
		a.n = (uint32_t)n;
		a.out_cap = (uint32_t)n;
		a.shift = h.shift[p];
		a.bits = h.bits[p];
		a.hist = s.hist + p * RADIX;
		a.status = s.status + (size_t)p * tiles * RADIX;
		a.ticket = s.ctl + p;
		launch_onesweep(a, n, stream);
		kin = a.kout;
		vin = a.vout;
	}
	return cudaGetLastError();
}

} // namespace brs
