// ASCII PLY rows of the reference's on-disk point clouds:
//   np.savetxt(f, hstack([points3D, colors]), '%f %f %f %d %d %d')        semantic_depth_lib/point_cloud_2_ply.py:62-70
// '%f' is the correctly rounded 6-decimal expansion of the value (round-half-even on the exact binary number).  A float32
// is m * 2^E with m < 2^24, so value * 10^6 = (m * 10^6) * 2^E is an exact 44-bit integer times a power of two: the rounding
// is decided in integer arithmetic (shift, remainder against half, tie to even), the integer part of huge values (up to
// 3.4e38) comes from a 128-bit shift.  No floating-point operation takes part in the formatting.
// Three kernels: row lengths -> per-tile totals; exclusive scan of the tile totals; formatting into the tile's slice of
// the output (staged in shared memory, written out with consecutive byte stores).
#include "sd_internal.cuh"

namespace sd {

constexpr int kPlyThreads = 256;
constexpr int kPlyStage = 20 * 1024;        // bytes of a tile staged in shared memory (typical rows are ~40 bytes)

// digits of v (most significant first) into dst (may be nullptr: count only); returns the number of digits (>= 1)
__device__ __forceinline__ int put_u64(char* dst, unsigned long long v) {
    char tmp[20]; int n = 0;
    do { tmp[n++] = (char)('0' + (int)(v % 10ull)); v /= 10ull; } while (v);
    if (dst) for (int i = 0; i < n; ++i) dst[i] = tmp[n - 1 - i];
    return n;
}
__device__ __noinline__ int put_u128(char* dst, unsigned __int128 v) {
    char tmp[40]; int n = 0;
    do { tmp[n++] = (char)('0' + (int)(v % 10)); v /= 10; } while (v);
    if (dst) for (int i = 0; i < n; ++i) dst[i] = tmp[n - 1 - i];
    return n;
}

// '%f' % value  (Python / C printf with the default precision 6)
__device__ __forceinline__ int fmt_f(char* dst, float f) {
    const uint32_t b = __float_as_uint(f);
    const bool neg = (b >> 31) != 0;
    const int e8 = (int)((b >> 23) & 0xffu);
    const uint32_t m23 = b & 0x7fffffu;
    int n = 0;
    if (e8 == 0xff) {                                         // 'inf' / '-inf' / 'nan' (Python prints nan without a sign)
        if (m23) { if (dst) { dst[0] = 'n'; dst[1] = 'a'; dst[2] = 'n'; } return 3; }
        if (neg) { if (dst) dst[0] = '-'; n = 1; }
        if (dst) { dst[n] = 'i'; dst[n + 1] = 'n'; dst[n + 2] = 'f'; }
        return n + 3;
    }
    if (neg) { if (dst) dst[0] = '-'; n = 1; }
    const unsigned long long m = e8 ? (unsigned long long)(m23 | 0x800000u) : (unsigned long long)m23;
    const int E = (e8 ? e8 : 1) - 150;                        // value = m * 2^E
    unsigned long long frac = 0ull;
    if (E >= 0) {
        if (E <= 40) n += put_u64(dst ? dst + n : nullptr, m << E);
        else n += put_u128(dst ? dst + n : nullptr, (unsigned __int128)m << E);
    } else {
        const unsigned long long P = m * 1000000ull;          // < 2^44
        const int s = -E;
        unsigned long long q = 0ull;
        if (s < 64) {
            q = P >> s;
            const unsigned long long rem = P & ((1ull << s) - 1ull), half = 1ull << (s - 1);
            if (rem > half || (rem == half && (q & 1ull))) ++q;
        }                                                      // s >= 64: P < 2^44 <= half, rounds to 0
        n += put_u64(dst ? dst + n : nullptr, q / 1000000ull);
        frac = q % 1000000ull;
    }
    if (dst) {
        dst[n] = '.';
        for (int i = 6; i >= 1; --i) { dst[n + i] = (char)('0' + (int)(frac % 10ull)); frac /= 10ull; }
    }
    return n + 7;
}
// the same for a float64: m < 2^53, value * 10^6 = (m * 10^6) * 2^E with m * 10^6 < 2^73 (128-bit integer arithmetic).
// Finite values of 2^128 and above are refused by the host wrapper (their integer part needs more than 128 bits).
__device__ __noinline__ int fmt_f(char* dst, double d) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(d);
    const bool neg = (b >> 63) != 0;
    const int e11 = (int)((b >> 52) & 0x7ffull);
    const unsigned long long m52 = b & 0xfffffffffffffull;
    int n = 0;
    if (e11 == 0x7ff) {
        if (m52) { if (dst) { dst[0] = 'n'; dst[1] = 'a'; dst[2] = 'n'; } return 3; }
        if (neg) { if (dst) dst[0] = '-'; n = 1; }
        if (dst) { dst[n] = 'i'; dst[n + 1] = 'n'; dst[n + 2] = 'f'; }
        return n + 3;
    }
    if (neg) { if (dst) dst[0] = '-'; n = 1; }
    const unsigned long long m = e11 ? (m52 | 0x10000000000000ull) : m52;
    const int E = (e11 ? e11 : 1) - 1075;                     // value = m * 2^E
    unsigned long long frac = 0ull;
    if (E >= 0) {
        n += put_u128(dst ? dst + n : nullptr, (unsigned __int128)m << min(E, 75));
    } else {
        const unsigned __int128 P = (unsigned __int128)m * 1000000ull;     // < 2^73
        const int s = -E;
        unsigned __int128 q = 0;
        if (s < 100) {
            q = P >> s;
            const unsigned __int128 one = 1;
            const unsigned __int128 rem = P & ((one << s) - 1), half = one << (s - 1);
            if (rem > half || (rem == half && ((unsigned long long)q & 1ull))) ++q;
        }                                                      // s >= 100: P < 2^73 < half, rounds to 0
        n += put_u64(dst ? dst + n : nullptr, (unsigned long long)(q / 1000000ull));      // < 2^53
        frac = (unsigned long long)(q % 1000000ull);
    }
    if (dst) {
        dst[n] = '.';
        for (int i = 6; i >= 1; --i) { dst[n + i] = (char)('0' + (int)(frac % 10ull)); frac /= 10ull; }
    }
    return n + 7;
}
__device__ __forceinline__ int fmt_u8(char* dst, unsigned v) {
    if (v >= 100u) { if (dst) { dst[0] = (char)('0' + v / 100u); dst[1] = (char)('0' + (v / 10u) % 10u); dst[2] = (char)('0' + v % 10u); } return 3; }
    if (v >= 10u) { if (dst) { dst[0] = (char)('0' + v / 10u); dst[1] = (char)('0' + v % 10u); } return 2; }
    if (dst) dst[0] = (char)('0' + v);
    return 1;
}
// "x y z r g b\n"
template <typename T>
__device__ __forceinline__ int fmt_row(char* dst, T x, T y, T z, unsigned r, unsigned g, unsigned b) {
    int n = fmt_f(dst, x);
    if (dst) dst[n] = ' '; ++n;
    n += fmt_f(dst ? dst + n : nullptr, y);
    if (dst) dst[n] = ' '; ++n;
    n += fmt_f(dst ? dst + n : nullptr, z);
    if (dst) dst[n] = ' '; ++n;
    n += fmt_u8(dst ? dst + n : nullptr, r);
    if (dst) dst[n] = ' '; ++n;
    n += fmt_u8(dst ? dst + n : nullptr, g);
    if (dst) dst[n] = ' '; ++n;
    n += fmt_u8(dst ? dst + n : nullptr, b);
    if (dst) dst[n] = '\n';
    return n + 1;
}

template <typename T>
struct PlyArgs {
    const T* x; const T* y; const T* z; const uint8_t* rgb; int n;
    char* out; unsigned long long capacity;
    uint32_t* tile_total; uint32_t* tile_off; unsigned long long* total; int ntiles;
};

template <typename T>
__global__ void __launch_bounds__(kPlyThreads)
ply_len_kernel(const PlyArgs<T> a) {
    __shared__ int s_scan[33];
    const int i = blockIdx.x * kPlyThreads + threadIdx.x;
    int len = 0;
    if (i < a.n) len = fmt_row(nullptr, __ldg(a.x + i), __ldg(a.y + i), __ldg(a.z + i), a.rgb[3 * i], a.rgb[3 * i + 1], a.rgb[3 * i + 2]);
    int total;
    block_excl_scan(len, s_scan, &total);
    if (threadIdx.x == 0) a.tile_total[blockIdx.x] = (uint32_t)total;
}

template <typename T>
__global__ void __launch_bounds__(1024)
ply_scan_kernel(const PlyArgs<T> a) {
    __shared__ int s_scan[33];
    unsigned long long run = 0ull;
    for (int base = 0; base < a.ntiles; base += 1024) {
        const int t = base + threadIdx.x;
        const int v = (t < a.ntiles) ? (int)a.tile_total[t] : 0;       // a chunk of 1024 tiles is < 2^31 bytes
        int total;
        const int excl = block_excl_scan(v, s_scan, &total);
        if (t < a.ntiles) a.tile_off[t] = (uint32_t)(run + (unsigned long long)excl);   // offsets fit 32 bits when the total does
        run += (unsigned long long)total;
    }
    if (threadIdx.x == 0) *a.total = run;
}

template <typename T>
__global__ void __launch_bounds__(kPlyThreads)
ply_write_kernel(const PlyArgs<T> a) {
    __shared__ int s_scan[33];
    __shared__ char s_buf[kPlyStage];
    if (*a.total > a.capacity) return;                         // the caller retries with a larger buffer
    const int i = blockIdx.x * kPlyThreads + threadIdx.x;
    T x = 0, y = 0, z = 0; unsigned r = 0, g = 0, b = 0;
    int len = 0;
    if (i < a.n) {
        x = __ldg(a.x + i); y = __ldg(a.y + i); z = __ldg(a.z + i);
        r = a.rgb[3 * i]; g = a.rgb[3 * i + 1]; b = a.rgb[3 * i + 2];
        len = fmt_row(nullptr, x, y, z, r, g, b);
    }
    int total;
    const int excl = block_excl_scan(len, s_scan, &total);
    char* tile_out = a.out + a.tile_off[blockIdx.x];
    if (total <= kPlyStage) {
        if (i < a.n) fmt_row(s_buf + excl, x, y, z, r, g, b);
        __syncthreads();
        for (int k = threadIdx.x; k < total; k += kPlyThreads) tile_out[k] = s_buf[k];
    } else if (i < a.n) {
        fmt_row(tile_out + excl, x, y, z, r, g, b);
    }
}

}  // namespace sd

template <typename T>
static int launch_ply_rows(const T* d_x, const T* d_y, const T* d_z, const uint8_t* d_rgb, int n, char* d_out,
                           unsigned long long capacity, uint32_t* d_tile_scratch, unsigned long long* d_total, cudaStream_t st) {
    using namespace sd;
    if (n <= 0) return SD_OK;
    PlyArgs<T> a;
    a.x = d_x; a.y = d_y; a.z = d_z; a.rgb = d_rgb; a.n = n; a.out = d_out; a.capacity = capacity;
    a.ntiles = ceil_div(n, kPlyThreads);
    a.tile_total = d_tile_scratch; a.tile_off = d_tile_scratch + a.ntiles; a.total = d_total;
    ply_len_kernel<T><<<a.ntiles, kPlyThreads, 0, st>>>(a);
    ply_scan_kernel<T><<<1, 1024, 0, st>>>(a);
    ply_write_kernel<T><<<a.ntiles, kPlyThreads, 0, st>>>(a);
    SD_LAUNCH_CHECK();
    return SD_OK;
}

int sd_launch_ply_rows(const float* d_x, const float* d_y, const float* d_z, const uint8_t* d_rgb, int n, char* d_out,
                       unsigned long long capacity, uint32_t* d_tile_scratch, unsigned long long* d_total, cudaStream_t st) {
    return launch_ply_rows<float>(d_x, d_y, d_z, d_rgb, n, d_out, capacity, d_tile_scratch, d_total, st);
}
int sd_launch_ply_rows_f64(const double* d_x, const double* d_y, const double* d_z, const uint8_t* d_rgb, int n, char* d_out,
                           unsigned long long capacity, uint32_t* d_tile_scratch, unsigned long long* d_total, cudaStream_t st) {
    return launch_ply_rows<double>(d_x, d_y, d_z, d_rgb, n, d_out, capacity, d_tile_scratch, d_total, st);
}
