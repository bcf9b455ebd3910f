/* plugin_oracle.c -- TEST INFRASTRUCTURE ONLY (the checker, never the product path).
 *
 * CPU restatement of two host-side loops of the plug-in itself that SURVEY.md section 8(f) lists as the rows next to
 * the hot path.  Unlike liblqr (oracle/lqr_oracle.c, parity unpinned), their source IS in the reference tree, so this
 * file follows it line by line in behaviour:
 *
 *   plugin_oracle_vmap_colour     write_vmap_to_layer, src/io_functions.c:249-279 -- the seam map (0 = never carved,
 *                                 k = k-th seam) drawn as RGBA: a lerp between two colours, alpha 0.5..1.
 *   plugin_oracle_guess_new_size  guess_new_size, src/layers_combo.c:274-392 -- the size left when every pixel of
 *                                 the discard mask goes: old size minus the largest per-line count of mask pixels.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may load this library.
 * Build: oracle/Makefile (gcc -O2 -ffp-contract=off: doubles are rounded exactly as written, as on x86-64 without FMA). */
#include <stddef.h>

#define PO_PUBLIC __attribute__((visibility("default")))

/* io_functions.c:249-279.  buffer: w*h seam orders (lqr_vmap_get_data), depth = lqr_vmap_get_depth; start / end:
 * the r, g, b of GimpRGB colour_start / colour_end (doubles in [0,1], render.c:341-342); out: w*h*4 bytes.
 * The double -> guchar assignments truncate (C conversion). */
PO_PUBLIC void plugin_oracle_vmap_colour(const int *buffer, int w, int h, int depth, const double start[3],
                                         const double end[3], unsigned char *out)
{
    const size_t n = (size_t) w * (size_t) h;
    for (size_t i = 0; i < n; i++) {
        unsigned char *px = out + 4 * i;
        const int vs = buffer[i];
        if (vs == 0) { /* io_functions.c:254-260 */
            px[0] = px[1] = px[2] = px[3] = 0;
            continue;
        }
        const double value = (double) (depth + 1 - vs) / (depth + 1); /* :263 */
        for (int k = 0; k < 3; k++) {
            const double c = value * start[k] + (1 - value) * end[k]; /* :264-266 */
            px[k] = (unsigned char) (255 * c);                         /* :268-270 */
        }
        const double al = 0.5 * (1 + value); /* :267 */
        px[3] = (unsigned char) (255 * al);  /* :271 */
    }
}

/* layers_combo.c:274-392.  mask: the discard layer, width x height x bpp, placed at (x_off, y_off) relative to the
 * layer being resized (old_width x old_height).  direction 0 = GUESS_DIR_HOR (lines are mask rows, returns the new
 * width), 1 = GUESS_DIR_VERT (lines are mask columns, returns the new height). */
PO_PUBLIC int plugin_oracle_guess_new_size(const unsigned char *mask, int width, int height, int bpp, int has_alpha,
                                           int x_off, int y_off, int old_width, int old_height, int direction)
{
    const int c_bpp = bpp - (has_alpha ? 1 : 0);                                 /* :313 */
    const int x_lo = x_off > 0 ? x_off : 0, y_lo = y_off > 0 ? y_off : 0;
    const int x_hi = old_width < width + x_off ? old_width : width + x_off;    /* :324-325 */
    const int y_hi = old_height < height + y_off ? old_height : height + y_off;
    const int lw = x_hi - x_lo, lh = y_hi - y_lo;
    const int old_size = direction == 0 ? old_width : old_height;               /* :296-303 */
    const int z1min = direction == 0 ? y_lo : x_lo;                             /* :329-338 */
    const int z1max = direction == 0 ? y_hi : x_hi;
    const int z2max = direction == 0 ? lw : lh;
    const int col0 = -x_off > 0 ? -x_off : 0, row0 = -y_off > 0 ? -y_off : 0; /* first mask column / row inside the layer */
    int max_mask_size = 0;
    for (int z1 = z1min; z1 < z1max; z1++) {
        int mask_size = 0;
        for (int z2 = 0; z2 < z2max; z2++) {
            /* :349-356: row (z1 - y_off) from column col0, or column (z1 - x_off) from row row0 */
            const unsigned char *px = direction == 0 ? mask + ((size_t) (z1 - y_off) * width + (size_t) (col0 + z2)) * bpp
                                                     : mask + ((size_t) (row0 + z2) * width + (size_t) (z1 - x_off)) * bpp;
            double sum = 0;
            for (int k = 0; k < c_bpp; k++) sum += px[k]; /* :361-365 */
            sum /= (255 * c_bpp);                          /* :367 */
            if (has_alpha) sum *= (double) px[bpp - 1] / 255; /* :368-371 */
            if (sum >= (0.5 / c_bpp)) mask_size++;         /* :373-376 */
        }
        if (mask_size > max_mask_size) max_mask_size = mask_size; /* :378-381 */
    }
    return old_size - max_mask_size; /* :385 */
}
