// Hardware probe (diagnostics, not on the product path): does a tcgen05.mma K-major SWIZZLE_128B operand
// descriptor accept a start address that is 128-byte- but not 1024-byte-aligned, and a stride between 8-row
// groups (SBO) other than 1024 B?  A "yes" lets a 3x3 convolution read all nine filter taps as shifted views
// of ONE halo tile in shared memory instead of nine im2col TMA loads.
//   A: [rows_a][64] bf16 (row-major) loaded with tiled TMA (128B swizzle) at a 1024-aligned smem base,
//   B: [64][64] bf16,  D[128][64] = A_view * B^T with A_view row r = A[shift + (r/8)*(sbo/128) + r%8].
#include "tc_common.cuh"

namespace sr {

__global__ void __launch_bounds__(128, 1)
umma_shift_probe_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, int rows_a,
                        int shift, int sbo_bytes, int base_offset, float* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sa = smem;                               // rows_a x 128 B
    uint8_t* sb = smem + (size_t)rows_a * 128;        // 64 x 128 B (rows_a is a multiple of 8)
    uint64_t* bars = reinterpret_cast<uint64_t*>(sb + 64 * 128);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(bars, 1); mbar_init(bars + 1, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot, 64);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) {
        mbar_expect_tx(bars, (uint32_t)(rows_a * 128 + 64 * 128));
        for (int r0 = 0; r0 < rows_a; r0 += 128) tma_load_2d(sa + (size_t)r0 * 128, &map_a, bars, 0, r0);
        tma_load_2d(sb, &map_b, bars, 0, 0);
        mbar_wait(bars, 0);
        tc_fence_after();
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((128u >> 4) << 24);
        uint64_t adesc = 0;
        const uint32_t a_addr = smem_u32(sa) + (uint32_t)shift * 128u;
        adesc |= (uint64_t)((a_addr & 0x3FFFF) >> 4);
        adesc |= (uint64_t)1 << 16;
        adesc |= (uint64_t)((uint32_t)sbo_bytes >> 4) << 32;
        adesc |= (uint64_t)1 << 46;
        adesc |= (uint64_t)(base_offset & 7) << 49;
        adesc |= (uint64_t)2 << 61;
        const uint64_t bdesc = make_kmajor_sw128_desc(smem_u32(sb));
        for (int k = 0; k < 4; ++k) umma_f16(tmem_base, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, k ? 1u : 0u);
        umma_commit(bars + 1);
    }
    __syncthreads();
    mbar_wait(bars + 1, 0);
    tc_fence_after();
    uint32_t v[32];
    for (int c0 = 0; c0 < 64; c0 += 32) {
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
        for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 64 + c0 + j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 0) tmem_dealloc(tmem_base, 64);
}

int debug_umma_shift(const void* a, int rows_a, const void* b, int shift, int sbo_bytes, int base_offset, float* out, cudaStream_t st) {
    int rc = load_driver_fns();
    if (rc != SR_OK) return rc;
    SR_REQUIRE(rows_a % 128 == 0 && rows_a >= 128 && rows_a <= 1536, "probe: rows_a must be a multiple of 128 in [128, 1536]");
    alignas(64) CUtensorMap map_a, map_b;
    rc = make_tiled2d_map(&map_a, a, (uint64_t)rows_a, 64, 128);
    if (rc != SR_OK) return rc;
    rc = make_tiled2d_map(&map_b, b, 64, 64, 64);
    if (rc != SR_OK) return rc;
    const size_t smem = 1024 + (size_t)rows_a * 128 + 64 * 128 + 64;
    cudaFuncSetAttribute(umma_shift_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    umma_shift_probe_kernel<<<1, 128, smem, st>>>(map_a, map_b, rows_a, shift, sbo_bytes, base_offset, out);
    count_launch();
    return check_launch("umma_shift_probe_kernel");
}

// ------------------------------------------------------------------------------------------------
// Issue-rate probe: one thread per CTA issues `iters` rounds of tcgen05.mma (M=128, N=n, K=16, bf16, both operands in
// shared memory, garbage data) round-robin over `num_acc` independent TMEM accumulators, then commits and waits.
// cycles[blockIdx.x] = clock64 ticks from the first issue to the completion of the last MMA.
// Answers: what is the cost of back-to-back MMAs that accumulate into the SAME accumulator vs different ones?
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1)
umma_rate_probe_kernel(int n, int num_acc, int iters, int k_steps, long long* cycles) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sa = smem;                      // 128 x 128 B
    uint8_t* sb = smem + 16384;              // 256 x 128 B
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 16384 + 32768);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 4);
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(bar + i, 1); fence_barrier_init(); }
    if (threadIdx.x < 32) tmem_alloc(tmem_slot, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // `issuers` = num_acc >> 8 (0 -> 1): that many warps issue concurrently (lane 0 of each), every issuer round-robins
    // over its own (num_acc & 0xff) accumulators; mode bit 16: four K-steps per asm statement
    const int issuers = ((num_acc >> 8) & 0xff) ? ((num_acc >> 8) & 0xff) : 1;
    const int x4 = (num_acc >> 16) & 1;
    const int nacc = num_acc & 0xff;
    const int warp = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0 && warp < issuers) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
        const uint64_t adesc = make_kmajor_sw128_desc(smem_u32(sa));
        const uint64_t bdesc = make_kmajor_sw128_desc(smem_u32(sb));
        const int cols = 512 / (issuers * nacc);
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            for (int a = 0; a < nacc; ++a) {
                const uint32_t d = tmem_base + (uint32_t)((warp * nacc + a) * cols);
                if (x4) { for (int k = 0; k < k_steps; k += 4) umma_f16_x4(d, adesc, bdesc, idesc, 1u); }
                else { for (int k = 0; k < k_steps; ++k) umma_f16(d, adesc + (uint64_t)((k & 3) * 2), bdesc + (uint64_t)((k & 3) * 2), idesc, 1u); }
            }
        }
        umma_commit(bar + warp);
        mbar_wait(bar + warp, 0);
        cycles[blockIdx.x * issuers + warp] = clock64() - t0;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (threadIdx.x < 32) tmem_dealloc(tmem_base, 512);
}

int debug_umma_rate(int n, int num_acc, int iters, int k_steps, int grid, long long* cycles, cudaStream_t st) {
    const int issuers = ((num_acc >> 8) & 0xff) ? ((num_acc >> 8) & 0xff) : 1, nacc = num_acc & 0xff;
    SR_REQUIRE(n >= 16 && n <= 256 && n % 16 == 0 && nacc >= 1 && issuers <= 4 && n * nacc * issuers <= 512 && iters > 0 && k_steps > 0 && grid > 0,
               "umma_rate probe: bad arguments");
    const size_t smem = 1024 + 16384 + 32768 + 64;
    cudaFuncSetAttribute(umma_rate_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    umma_rate_probe_kernel<<<grid, 128, smem, st>>>(n, num_acc, iters, k_steps, cycles);
    count_launch();
    return check_launch("umma_rate_probe_kernel");
}

// Store-pattern probe: the epilogue of a convolution writes [rows][row_bytes] bf16 outputs of which one warp owns a 32-row x
// 64-byte chunk at a time.  pattern 0: lane = row, four 16-byte stores (32 lines per instruction); 1: lane quads share a row
// (64 contiguous bytes per row, 8 rows per instruction); 2: lane octets write 128 bytes of a row (4 rows per instruction, chunk =
// 128 bytes); 3: the whole warp writes 512 contiguous bytes (1 row per instruction, chunk = 512 bytes).  Every pattern writes the
// same `rows * row_bytes` bytes once; the time per launch tells which of L2 write bandwidth / LSU line rate bounds the epilogue.
__global__ void __launch_bounds__(256, 1)
store_pattern_probe_kernel(uint4* out, int rows, int row_bytes, int pattern) {
    const int warp = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31, warps = gridDim.x * 8;
    const int units_per_row = row_bytes / 16;                      // 16-byte units
    const uint4 v = make_uint4(warp, lane, 0x3f803f80u, 0x3f803f80u);
    if (pattern == 0 || pattern == 1) {
        const int chunks_per_row = row_bytes / 64;
        const int tiles = (rows / 32) * chunks_per_row;
        for (int t = warp; t < tiles; t += warps) {
            const int rb = (t / chunks_per_row) * 32, cu = (t % chunks_per_row) * 4;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int row = pattern == 0 ? rb + lane : rb + (lane & ~3) + i;
                const int unit = pattern == 0 ? cu + i : cu + (lane & 3);
                out[(size_t)row * units_per_row + unit] = v;
            }
        }
    } else if (pattern == 2) {
        const int chunks_per_row = row_bytes / 128;
        const int tiles = (rows / 32) * chunks_per_row;
        for (int t = warp; t < tiles; t += warps) {
            const int rb = (t / chunks_per_row) * 32, cu = (t % chunks_per_row) * 8;
#pragma unroll
            for (int i = 0; i < 8; ++i) out[(size_t)(rb + i * 4 + (lane >> 3)) * units_per_row + cu + (lane & 7)] = v;
        }
    } else {
        const size_t total = (size_t)rows * units_per_row;
        for (size_t u = (size_t)warp * 32 + lane; u < total; u += (size_t)warps * 32) out[u] = v;
    }
}

int debug_store_pattern(void* out, int rows, int row_bytes, int pattern, int grid, cudaStream_t st) {
    SR_REQUIRE(out && rows % 32 == 0 && row_bytes % 512 == 0 && pattern >= 0 && pattern <= 3 && grid > 0, "store_pattern probe: bad arguments");
    store_pattern_probe_kernel<<<grid, 256, 0, st>>>((uint4*)out, rows, row_bytes, pattern);
    count_launch();
    return check_launch("store_pattern_probe_kernel");
}

}  // namespace sr
