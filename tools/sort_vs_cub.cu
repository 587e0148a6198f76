// brs_sort_pairs_u32 (libbloomrast's hand-written stable radix sort) against
// cub::DeviceRadixSort::SortPairs<uint32_t, uint32_t> on the sub-problem the pipeline actually has:
// n = 100 K / 1 M / 3 M (key, id) pairs, 24 significant key bits (depth bits minus the bias).
// SURVEY.md §2.2 names CUB's sm_100 onesweep as the kernel to beat (reference rasterizer_impl.cu:304-309).
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/sort_vs_cub.cu -Iinclude \
//        -Lbloomscene_b200 -lbloomrast -Xlinker -rpath -Xlinker '$ORIGIN/..' -o bloomscene_b200/_build/sort_vs_cub
//
// Keys stay L2-resident between iterations on purpose: in the pipeline the keys were written by the
// previous kernel a few microseconds earlier.  Output: one JSON line per (n, distribution).
#include <cub/cub.cuh>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "bloomrast.h"

#define CK(x)                                                                                                          \
	do {                                                                                                               \
		cudaError_t e_ = (x);                                                                                          \
		if (e_ != cudaSuccess) {                                                                                       \
			fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_));                                 \
			exit(1);                                                                                                   \
		}                                                                                                              \
	} while (0)

static float median(std::vector<float> v)
{
	std::sort(v.begin(), v.end());
	return v[v.size() / 2];
}

int main(int argc, char** argv)
{
	const int iters = argc > 1 ? atoi(argv[1]) : 50;
	const int sizes[] = {50000, 100000, 250000, 500000, 1000000, 3000000};
	const char* dists[] = {"uniform24", "depth"};
	cudaStream_t stream;
	CK(cudaStreamCreate(&stream));
	cudaEvent_t e0, e1;
	CK(cudaEventCreate(&e0));
	CK(cudaEventCreate(&e1));
	for (int n : sizes) {
		for (int di = 0; di < 2; di++) {
			std::mt19937 rng(1234 + n + di);
			std::vector<uint32_t> h(n);
			uint32_t span = 0;
			if (di == 0) {
				for (auto& k : h)
					k = rng() & 0xFFFFFFu;
			} else {
				// view-space depths of the "object" scene: z in [2.2, 4.2]; key = bits(z) - (min bits & ~255)
				std::uniform_real_distribution<float> u(2.2f, 4.2f);
				uint32_t mn = 0xFFFFFFFFu;
				for (auto& k : h) {
					float z = u(rng);
					memcpy(&k, &z, 4);
					mn = std::min(mn, k);
				}
				const uint32_t bias = mn & ~0xFFu;
				for (auto& k : h)
					k -= bias;
			}
			for (auto k : h)
				span = std::max(span, k);
			int end_bit = 1;
			while (end_bit < 32 && (span >> end_bit))
				end_bit++;

			uint32_t *kin, *kout_a, *vout_a, *kout_b, *vout_b, *vin;
			CK(cudaMalloc(&kin, 4ull * n));
			CK(cudaMalloc(&vin, 4ull * n));
			CK(cudaMalloc(&kout_a, 4ull * n));
			CK(cudaMalloc(&vout_a, 4ull * n));
			CK(cudaMalloc(&kout_b, 4ull * n));
			CK(cudaMalloc(&vout_b, 4ull * n));
			CK(cudaMemcpy(kin, h.data(), 4ull * n, cudaMemcpyHostToDevice));
			std::vector<uint32_t> iota(n);
			for (int i = 0; i < n; i++)
				iota[i] = i;
			CK(cudaMemcpy(vin, iota.data(), 4ull * n, cudaMemcpyHostToDevice));

			void* scratch;
			CK(cudaMalloc(&scratch, brs_sort_scratch_bytes(n)));
			size_t cub_bytes = 0;
			CK(cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, kin, kout_b, vin, vout_b, n, 0, end_bit, stream));
			void* cub_tmp;
			CK(cudaMalloc(&cub_tmp, cub_bytes));

			std::vector<float> t_ours, t_cub;
			for (int it = 0; it < iters + 5; it++) {
				CK(cudaEventRecord(e0, stream));
				int st = brs_sort_pairs_u32(kin, vin, kout_a, vout_a, n, 0, end_bit, scratch, (brs_stream)stream);
				CK(cudaEventRecord(e1, stream));
				CK(cudaEventSynchronize(e1));
				if (st != 0) {
					fprintf(stderr, "brs_sort_pairs_u32: %s\n", brs_error_string(st));
					return 1;
				}
				float ms;
				CK(cudaEventElapsedTime(&ms, e0, e1));
				if (it >= 5)
					t_ours.push_back(ms);
				CK(cudaEventRecord(e0, stream));
				CK(cub::DeviceRadixSort::SortPairs(cub_tmp, cub_bytes, kin, kout_b, vin, vout_b, n, 0, end_bit, stream));
				CK(cudaEventRecord(e1, stream));
				CK(cudaEventSynchronize(e1));
				CK(cudaEventElapsedTime(&ms, e0, e1));
				if (it >= 5)
					t_cub.push_back(ms);
			}
			std::vector<uint32_t> ka(n), va(n), kb(n), vb(n);
			CK(cudaMemcpy(ka.data(), kout_a, 4ull * n, cudaMemcpyDeviceToHost));
			CK(cudaMemcpy(va.data(), vout_a, 4ull * n, cudaMemcpyDeviceToHost));
			CK(cudaMemcpy(kb.data(), kout_b, 4ull * n, cudaMemcpyDeviceToHost));
			CK(cudaMemcpy(vb.data(), vout_b, 4ull * n, cudaMemcpyDeviceToHost));
			const bool same = ka == kb && va == vb;
			const float mo = median(t_ours), mc = median(t_cub);
			printf("{\"n\": %d, \"keys\": \"%s\", \"key_bits\": %d, \"iters\": %d, \"brs_sort_pairs_u32_ms\": %.4f, "
			       "\"cub_SortPairs_ms\": %.4f, \"ours_over_cub\": %.3f, \"identical_output\": %s, "
			       "\"ours_GBps_8B_pairs_rw_per_pass\": %.1f}\n",
			       n, dists[di], end_bit, iters, mo, mc, mo / mc, same ? "true" : "false",
			       16.0 * n * ((end_bit + 7) / 8) / (mo * 1e-3) / 1e9);
			fflush(stdout);
			cudaFree(kin); cudaFree(vin); cudaFree(kout_a); cudaFree(vout_a); cudaFree(kout_b); cudaFree(vout_b);
			cudaFree(scratch); cudaFree(cub_tmp);
		}
	}
	return 0;
}
