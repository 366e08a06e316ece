// Host-side exhaustive-ish check of the FastDiv magic numbers used by the persistent kernels' tile decode
// (lifelong-nnunet_b200/csrc/tc_common.cuh).  fd_div itself is a __device__ function; the host mirror below is the same
// expression with the 32x32->64 multiply written out.  Built and run by tests/test_abi.py::test_fastdiv_exact.
#include <cstdint>
#include <cstdio>
#include "../../lifelong-nnunet_b200/csrc/tc_common.cuh"

static int host_fd_div(int n, const b2::FastDiv& f) {
    if (f.d == 1) return n;
    const uint32_t hi = (uint32_t)(((uint64_t)(uint32_t)n * (uint64_t)f.m) >> 32);   // __umulhi
    return (int)(hi >> f.sh);
}

int main() {
    uint64_t rng = 0x9E3779B97F4A7C15ull;
    long long checked = 0;
    for (int d = 1; d <= 70000; d = d < 5000 ? d + 1 : d + 997) {
        const b2::FastDiv f = b2::make_fastdiv(d);
        // small n exhaustively, boundaries around multiples of d, random n up to 2^31 - 1
        for (int n = 0; n < 3000; ++n, ++checked)
            if (host_fd_div(n, f) != n / d) { printf("FAIL d=%d n=%d\n", d, n); return 1; }
        for (int k = 1; k < 200; ++k)
            for (int e = -1; e <= 1; ++e) {
                const long long n = (long long)k * 10007 * d + e;
                if (n < 0 || n > 0x7fffffffLL) continue;
                ++checked;
                if (host_fd_div((int)n, f) != (int)(n / d)) { printf("FAIL d=%d n=%lld\n", d, n); return 1; }
            }
        for (int r = 0; r < 2000; ++r, ++checked) {
            rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17;
            const int n = (int)(rng & 0x7fffffff);
            if (host_fd_div(n, f) != n / d) { printf("FAIL d=%d n=%d\n", d, n); return 1; }
        }
        const int nmax = 0x7fffffff;
        if (host_fd_div(nmax, f) != nmax / d) { printf("FAIL d=%d n=max\n", d); return 1; }
    }
    printf("OK %lld checks\n", checked);
    return 0;
}
