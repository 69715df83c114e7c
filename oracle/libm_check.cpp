/*
 * libm_check.cpp -- TEST INFRASTRUCTURE ONLY.  Compares the device models of cosf / sinf / atanf / atan2f in
 * spl_slam_b200/csrc/plf_libm.cuh (compiled here as host code, no FMA contraction) with this machine's libm, bit for bit.
 *
 *   libm_check quick   strided sample, a few seconds (run by tests/test_oracle_vs_ref.py)
 *   libm_check full    cosf/sinf over EVERY float with |x| <= 6.3, atanf over every finite float, atan2f on 1.6e9 pairs
 *
 * Build: g++ -O2 -ffp-contract=off -fopenmp oracle/libm_check.cpp -o oracle/_ref/libm_check -lm
 * Prints one line per function and exits non-zero on any mismatch.
 */
#include "../spl_slam_b200/csrc/plf_libm.cuh"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace plf_libm;

int main(int argc, char** argv)
{
    const bool full = argc > 1 && !strcmp(argv[1], "full");
    const uint32_t stride = full ? 1 : 257;
    const uint32_t hi = asuint(6.3f);
    long badc = 0, bads = 0, bada = 0, bad2 = 0, n1 = 0, n2 = 0, n3 = 0;
#pragma omp parallel for reduction(+ : badc, bads, n1) schedule(static)
    for (uint32_t u = 0; u <= hi; u += stride)
        for (int sgn = 0; sgn < 2; sgn++) {
            const float x = asfloat(u | ((uint32_t)sgn << 31));
            if (asuint(cosf(x)) != asuint(cosf_glibc(x))) badc++;
            if (asuint(sinf(x)) != asuint(sinf_glibc(x))) bads++;
            n1++;
        }
    printf("cosf: %ld mismatches, sinf: %ld mismatches over %ld floats with |x| <= 6.3\n", badc, bads, n1);
#pragma omp parallel for reduction(+ : bada, n2) schedule(static)
    for (uint32_t u = 0; u < 0x7f800000u; u += stride)
        for (int sgn = 0; sgn < 2; sgn++) {
            const float x = asfloat(u | ((uint32_t)sgn << 31));
            if (asuint(atanf(x)) != asuint(atanf_glibc(x))) bada++;
            n2++;
        }
    printf("atanf: %ld mismatches over %ld finite floats\n", bada, n2);
    const long per_thread = full ? 200000000L : 2000000L;
#pragma omp parallel reduction(+ : bad2, n3)
    {
        unsigned seed = 1234;
#ifdef _OPENMP
        seed += (unsigned)omp_get_thread_num();
#endif
        for (long i = 0; i < per_thread; i++) {
            /* coordinate differences as the reference forms them: multiples of small powers of two and general floats */
            float y = (float)((int)(rand_r(&seed) % 4000001) - 2000000) / 1024.0f * ((rand_r(&seed) & 1) ? 1.0f : 0.37f);
            float x = (float)((int)(rand_r(&seed) % 4000001) - 2000000) / 1024.0f;
            if ((i & 1023) == 0) x = 0.0f;
            if ((i & 1023) == 1) y = 0.0f;
            if ((i & 1023) == 2) x = 1.0f;
            if (asuint(atan2f(y, x)) != asuint(atan2f_glibc(y, x))) bad2++;
            n3++;
        }
    }
    printf("atan2f: %ld mismatches over %ld pairs\n", bad2, n3);
    return (badc || bads || bada || bad2) ? 1 : 0;
}
