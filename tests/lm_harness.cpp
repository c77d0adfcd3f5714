// TEST INFRASTRUCTURE ONLY -- runs tb_lm_fit_warp (tobias_b200/csrc/tb_lm.cuh) inside the kernel emulator (one warp =
// 32 fibers) on windows read from a text file, so that tests/test_lm_exact.py can compare it with scipy.optimize.curve_fit.
// usage: lm_harness <file> <seq 0|1|2>   (0 warp-tree sums, 1 warp with MINPACK-order sums, 2 thread-per-window exact, 4 thread-per-window one-pass form, 5 the same with the fused trial + next head)     file: repeated blocks, a line "W <m>" followed by m lines "<x> <y>"
// prints one line per window: a b info nfev usable   (%.17g)
#define TB_EMU 1
#define TB_FIRST_GEN_FITS 1
#include "tb_common.cuh"
#include "tb_lm.cuh"
#include "tb_lm_thread.cuh"
#define TB_FIT_Q 64
#include <vector>
struct fake_ctx { long launches; };
static double gx[128], gy[128];
static int gm, g_seq;
static tb_lm_result gres;
static double gtx[10 * TB_FIT_Q];
static unsigned gty[10 * TB_FIT_Q];
void kern() {
    int lane = threadIdx.x & 31;
    if (g_seq == 2 || g_seq == 4 || g_seq == 5) {                    // thread-per-window form (tb_lm_thread.cuh): lane 0 fits the window alone
        if (lane == 0) {
            for (int i = 0; i < gm; i++) {
                gtx[(i % 10) * TB_FIT_Q + i / 10] = gx[i];
                gty[(i % 10) * TB_FIT_Q + i / 10] = (unsigned)gy[i];
            }
            tb_win_ref W{gtx, gty, TB_FIT_Q};
            if (g_seq == 2) tb_lm_fit_thread<false>(W, gm, gres); else if (g_seq == 4) tb_lm_fit_thread<true>(W, gm, gres); else tb_lm_fit_thread<true, true>(W, gm, gres);
        }
        return;
    }
    double x[4], y[4];
    for (int q = 0; q < 4; q++) { int i = lane + 32 * q; x[q] = i < gm ? gx[i] : 0; y[q] = i < gm ? gy[i] : 0; }
    static double wbuf[128];
    tb_lm_result r;
    if (g_seq) tb_lm_fit_warp<4, true>(x, y, gm, lane, wbuf, r); else tb_lm_fit_warp<4, false>(x, y, gm, lane, wbuf, r);
    if (lane == 0) gres = r;
}
int main(int argc, char **argv) {
    if (argc < 3) return 2;
    g_seq = atoi(argv[2]);
    FILE *f = fopen(argv[1], "r");
    if (!f) return 3;
    char tag[8];
    fake_ctx c{0};
    while (fscanf(f, "%7s %d", tag, &gm) == 2) {
        if (gm < 2 || gm > 128) return 4;
        for (int i = 0; i < gm; i++) if (fscanf(f, "%lf %lf", &gx[i], &gy[i]) != 2) return 5;
        TB_LAUNCH(&c, kern, 1, 32, 0);
        printf("%.17g %.17g %d %d %d\n", gres.a, gres.b, gres.info, gres.nfev, gres.usable);
    }
    fclose(f);
    return 0;
}
