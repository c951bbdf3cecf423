/* TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's monotonic alignment
 * search.  Follows /root/reference/TTS/tts/utils/monotonic_align/core.pyx:11-37
 * (maximum_path_each) and :42-47 (maximum_path_c, serial over the batch because the
 * reference is built without OpenMP, setup.py:73-92).
 *
 * Pinned against the compiled reference kernel (oracle/_ref/core*.so) by
 * tests/test_oracle_mas.py and against tests/golden/mas_*.npz.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The product path never does.
 *
 * One deliberate deviation: when t_x > t_y the reference's backtrack evaluates
 * value[index, -1] at y == 0 (bounds checks disabled, core.pyx:36); the result is
 * never used because the loop ends, so this port skips that read.
 */
#include <stdint.h>

static void mas_each(int32_t *path, float *value, int t_x, int t_y, int ld, float max_neg_val)
{
    int index = t_x - 1;
    for (int y = 0; y < t_y; ++y) {
        int lo = t_x + y - t_y; if (lo < 0) lo = 0;
        int hi = (t_x < y + 1) ? t_x : (y + 1);
        for (int x = lo; x < hi; ++x) {
            float v_cur = (x == y) ? max_neg_val : value[(long)x * ld + (y - 1)];
            float v_prev;
            if (x == 0) v_prev = (y == 0) ? 0.f : max_neg_val;
            else        v_prev = value[(long)(x - 1) * ld + (y - 1)];
            float m = (v_cur > v_prev) ? v_cur : v_prev;   /* C max() of the Cython source */
            value[(long)x * ld + y] = m + value[(long)x * ld + y];
        }
    }
    for (int y = t_y - 1; y >= 0; --y) {
        path[(long)index * ld + y] = 1;
        if (index != 0 && y > 0 &&
            (index == y || value[(long)index * ld + (y - 1)] < value[(long)(index - 1) * ld + (y - 1)]))
            index -= 1;
    }
}

/* values [B,Tx,Ty] (modified in place like the reference), paths [B,Tx,Ty] pre-zeroed */
void mas_oracle_f32(int32_t *paths, float *values, const int32_t *t_xs, const int32_t *t_ys,
                    int B, int Tx, int Ty, float max_neg_val)
{
    for (int b = 0; b < B; ++b)
        mas_each(paths + (long)b * Tx * Ty, values + (long)b * Tx * Ty, t_xs[b], t_ys[b], Ty, max_neg_val);
}
