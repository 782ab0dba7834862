/* Exhaustive proof-by-enumeration that the 3-instruction division used in the fused kernel,
 *     y = RN(1/b);  q0 = RN(a*y);  r = fma(-b, q0, a);  q = fma(r, y, q0)
 * equals the IEEE quotient RN(a/b) for every float a in [0, amax] outside the deep-subnormal range
 * (mismatches exist only where r or q underflows, a < 2^-102; the kernel's a = fl(x + r) is 0 or >= 2^-24) for a divisor b
 * (b = 2*bev_range; lidar_agent.py:548 divides by 2*r).  Build: gcc -O2 -mfma -ffp-contract=off.
 * usage: markstein_check <b> <amax>   -> prints mismatches (0 expected) */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
int main(int argc, char** argv) {
    float b = argc > 1 ? strtof(argv[1], 0) : 100.0f;
    float amax = argc > 2 ? strtof(argv[2], 0) : 100.0f;
    float y = 1.0f / b;
    uint32_t hi; memcpy(&hi, &amax, 4);
    uint64_t bad = 0, n = 0; float worst = 0.0f;
    for (uint32_t u = 0; u <= hi; ++u) {
        float a; memcpy(&a, &u, 4);
        float q0 = a * y;
        float r = fmaf(-b, q0, a);
        float q = fmaf(r, y, q0);
        float ref = a / b;
        if (q != ref) { if (a > worst) worst = a; ++bad; }
        ++n;
    }
    printf("b=%g amax=%g checked=%llu mismatches=%llu largest_mismatching_a=%a (%g)\n", b, amax, (unsigned long long)n, (unsigned long long)bad, worst, worst);
    return worst >= 1.1754944e-38f * 16777216.0f; /* fail only if a mismatch exists at or above 2^-102 */
}
