/* lqr_oracle.c -- CPU ORACLE for the seam-carving hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * What this is
 *   A single-threaded plain-C restatement of the algorithm that the external library
 *   liblqr (lqr-1 >= 0.4.0, Windows bundle pins 0.4.1: reference configure.ac:67-70,
 *   windows_installer_files/lqr-pack4win/winpack.sh:8) runs behind gimp-lqr-plugin's
 *   render path (reference src/render.c:222-248,318,366,376 and src/io_functions.c:94,125,155).
 *   liblqr is NOT vendored in the reference tree and is not installed in this image, so its
 *   published algorithm is restated here from SURVEY.md Appendix A (sections cited as A.n
 *   below) and anchored on the reference's own call sites (cited as file:line).
 *
 * PARITY UNPINNED
 *   The reference ships no tests, fixtures or golden vectors for this path and liblqr cannot
 *   be built here, so "oracle == liblqr" is unverified.  What the tests can and do pin is
 *   (a) the documented properties of the path (help/en/index.wiki:71,82,126,130), (b) hand
 *   computable micro cases, (c) CUDA engine == this oracle, bit for bit.
 *
 * Who may use it
 *   Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg.
 *   The product (liblqr-1.so -> libb200carve.so) never links, loads or calls this file.
 *
 * Build:  gcc -O2 -ffp-contract=off -fPIC -shared -Iinclude oracle/lqr_oracle.c -lm -o oracle/liblqr_oracle.so
 *   (-ffp-contract=off: liblqr's x86-64 builds have no FMA contraction; the CUDA kernels use
 *    __fadd_rn/__fmul_rn/__dadd_rn/__dmul_rn on the same expressions.)
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "lqr.h"

#define OMIN(a, b) ((a) < (b) ? (a) : (b))
#define OMAX(a, b) ((a) > (b) ? (a) : (b))
#define UPDATE_TOLERANCE (1e-5) /* A.8 */

enum { READ_BRIGHTNESS = 0, READ_LUMA = 1 };
enum { GRAD_NORM = 0, GRAD_SUMABS = 1, GRAD_XABS = 2, GRAD_NULL = 3 };

struct _LqrProgress {
    gfloat update_step;
    LqrProgressFuncInit init;
    LqrProgressFuncUpdate update;
    LqrProgressFuncEnd end;
    gchar init_width_message[LQR_PROGRESS_MAX_MESSAGE_LENGTH];
    gchar end_width_message[LQR_PROGRESS_MAX_MESSAGE_LENGTH];
    gchar init_height_message[LQR_PROGRESS_MAX_MESSAGE_LENGTH];
    gchar end_height_message[LQR_PROGRESS_MAX_MESSAGE_LENGTH];
};

struct _LqrVMap {
    gint *buffer;
    gint width, height, depth, orientation;
};

struct _LqrVMapList {
    LqrVMap *current;
    LqrVMapList *next;
};

struct _LqrCarverList {
    LqrCarver *current;
    LqrCarverList *next;
};

/* A.1 state */
struct _LqrCarver {
    gint w_start, h_start; /* reference size */
    gint w, h;             /* current size */
    gint w0, h0;           /* allocated map size */
    gint level, max_level;
    gint channels, alpha_channel;
    gint transposed;
    gboolean active, nrg_active;
    LqrCarver *root;
    LqrCarverList *attached;
    LqrVMapList *flushed_vs;
    gboolean dump_vmaps;
    LqrResizeOrder resize_order;
    LqrProgress *progress;
    gint session_update_step, session_rescale_total, session_rescale_current;

    guchar *rgb;
    gint *vs;
    gfloat *en, *bias, *m, *rigmask;
    gint *least;
    gint *raw_store;
    gint **raw;
    gint *vpath, *vpath_x, *nrg_xmin, *nrg_xmax;
    gdouble *rcache;

    gfloat rigidity;
    gfloat *rigmap_store; /* 2*delta_x+1 */
    gfloat *rigmap;       /* centred */
    gint delta_x;

    gint ef_index, grad_kind, read_kind, nrg_radius;
    gboolean nrg_uptodate;
    gint leftright;
    guint lr_switch_frequency;
    gfloat enl_step;

    /* read-out cursor (A.12) */
    gint cur_x, cur_y, cur_now;
    gboolean cur_eoc;
    guchar *line;

    /* instrumentation for tests / kernel design (not part of liblqr) */
    long stat_update_rows, stat_update_cells, stat_update_maxband;
    long stat_band_hist[16]; /* rows whose band width w satisfies 2^(i-1) < w <= 2^i (i = 0: w <= 1) */
};

/* ------------------------------------------------------------------ progress (A.10) */
LqrProgress *lqr_progress_new(void)
{
    LqrProgress *p = (LqrProgress *) calloc(1, sizeof(LqrProgress));
    if (!p) return NULL;
    p->update_step = 0.02f;
    strcpy(p->init_width_message, "Resizing width...");
    strcpy(p->end_width_message, "done");
    strcpy(p->init_height_message, "Resizing height...");
    strcpy(p->end_height_message, "done");
    return p;
}
LqrRetVal lqr_progress_set_init(LqrProgress *p, LqrProgressFuncInit f) { if (!p) return LQR_ERROR; p->init = f; return LQR_OK; }
LqrRetVal lqr_progress_set_update(LqrProgress *p, LqrProgressFuncUpdate f) { if (!p) return LQR_ERROR; p->update = f; return LQR_OK; }
LqrRetVal lqr_progress_set_end(LqrProgress *p, LqrProgressFuncEnd f) { if (!p) return LQR_ERROR; p->end = f; return LQR_OK; }
LqrRetVal lqr_progress_set_update_step(LqrProgress *p, gfloat s) { if (!p) return LQR_ERROR; p->update_step = s; return LQR_OK; }
static LqrRetVal set_msg(gchar *dst, const gchar *src)
{
    if (!src) return LQR_ERROR;
    strncpy(dst, src, LQR_PROGRESS_MAX_MESSAGE_LENGTH - 1);
    dst[LQR_PROGRESS_MAX_MESSAGE_LENGTH - 1] = 0;
    return LQR_OK;
}
LqrRetVal lqr_progress_set_init_width_message(LqrProgress *p, const gchar *m) { return p ? set_msg(p->init_width_message, m) : LQR_ERROR; }
LqrRetVal lqr_progress_set_init_height_message(LqrProgress *p, const gchar *m) { return p ? set_msg(p->init_height_message, m) : LQR_ERROR; }
LqrRetVal lqr_progress_set_end_width_message(LqrProgress *p, const gchar *m) { return p ? set_msg(p->end_width_message, m) : LQR_ERROR; }
LqrRetVal lqr_progress_set_end_height_message(LqrProgress *p, const gchar *m) { return p ? set_msg(p->end_height_message, m) : LQR_ERROR; }

static LqrRetVal progress_init(LqrProgress *p, const gchar *msg) { return (p && p->init) ? p->init(msg) : LQR_OK; }
static LqrRetVal progress_update(LqrProgress *p, gdouble f) { return (p && p->update) ? p->update(f) : LQR_OK; }
static LqrRetVal progress_end(LqrProgress *p, const gchar *msg) { return (p && p->end) ? p->end(msg) : LQR_OK; }

/* ------------------------------------------------------------------ cursor (A.12) */
static int invisible(const LqrCarver *r, gint z) { return r->vs[z] != 0 && r->vs[z] < r->level; }

static void cursor_reset(LqrCarver *r)
{
    r->cur_x = 0;
    r->cur_y = 0;
    r->cur_now = 0;
    r->cur_eoc = FALSE;
    while (invisible(r, r->cur_now)) r->cur_now++;
}

static void cursor_next(LqrCarver *r)
{
    if (r->cur_eoc) return;
    if (r->cur_x == r->w - 1) {
        if (r->cur_y == r->h - 1) {
            r->cur_eoc = TRUE;
            return;
        }
        r->cur_x = 0;
        r->cur_y++;
    } else {
        r->cur_x++;
    }
    r->cur_now++;
    while (invisible(r, r->cur_now)) r->cur_now++;
}

static gint cursor_left(const LqrCarver *r)
{
    gint z = r->cur_now - 1;
    while (invisible(r, z)) z--;
    return z;
}

/* ------------------------------------------------------------------ small helpers */
static void set_width(LqrCarver *r, gint w1)
{
    r->w = w1;
    r->level = r->w0 - w1 + 1;
}

typedef LqrRetVal (*CarverFn)(LqrCarver *, gint);

static LqrRetVal foreach_attached(LqrCarver *r, CarverFn fn, gint arg)
{
    LqrCarverList *it;
    for (it = r->attached; it; it = it->next) LQR_CATCH(fn(it->current, arg));
    return LQR_OK;
}

static LqrRetVal set_width_attached(LqrCarver *r, gint w1)
{
    set_width(r, w1);
    return foreach_attached(r, set_width_attached, w1);
}

static void propagate_vs(LqrCarver *r)
{
    LqrCarverList *it;
    for (it = r->attached; it; it = it->next) {
        it->current->vs = r->vs;
        propagate_vs(it->current);
    }
}

/* ------------------------------------------------------------------ pixel reading (A.2) */
static gdouble norm8(const guchar *rgb, gint idx) { return (gdouble) rgb[idx] / 255.0; }

static gdouble read_pixel_scalar(const LqrCarver *r, gint z)
{
    const guchar *p = r->rgb;
    gint c = r->channels;
    gdouble v;
    if (c <= 2) {
        v = norm8(p, z * c);
    } else {
        gdouble red = norm8(p, z * c), green = norm8(p, z * c + 1), blue = norm8(p, z * c + 2);
        if (r->read_kind == READ_LUMA) {
            v = 0.2126 * red + 0.7152 * green + 0.0722 * blue;
        } else {
            v = (red + green + blue) / 3;
        }
    }
    if (r->alpha_channel >= 0) v = v * norm8(p, z * c + r->alpha_channel);
    return v;
}

static LqrRetVal build_rcache(LqrCarver *r)
{
    gint x, y;
    r->rcache = (gdouble *) malloc(sizeof(gdouble) * (size_t) r->w0 * r->h0);
    LQR_CATCH_MEM(r->rcache);
    for (y = 0; y < r->h; y++)
        for (x = 0; x < r->w; x++) {
            gint z = r->raw[y][x];
            r->rcache[z] = read_pixel_scalar(r, z);
        }
    return LQR_OK;
}

/* ------------------------------------------------------------------ energy (A.3) */
static gfloat grad_eval(gint kind, gdouble gx, gdouble gy)
{
    switch (kind) {
        case GRAD_NORM: return (gfloat) sqrt(gx * gx + gy * gy);
        case GRAD_SUMABS: return (gfloat) ((fabs(gx) + fabs(gy)) / 2);
        case GRAD_XABS: return (gfloat) fabs(gx);
        default: return 0;
    }
}

static void compute_e(LqrCarver *r, gint x, gint y)
{
    gint z = r->raw[y][x];
    gfloat e = 0, b_add = 0;
    if (r->grad_kind != GRAD_NULL) {
        const gdouble *b = r->rcache;
        gint **raw = r->raw;
        gdouble gx, gy;
        /* reads that fall outside the image return 0 (only reachable when a dimension is 1) */
        if (y == 0) gy = (r->h > 1 ? b[raw[y + 1][x]] : 0) - b[z];
        else if (y < r->h - 1) gy = (b[raw[y + 1][x]] - b[raw[y - 1][x]]) / 2;
        else gy = b[z] - b[raw[y - 1][x]];
        if (x == 0) gx = (r->w > 1 ? b[raw[y][x + 1]] : 0) - b[z];
        else if (x < r->w - 1) gx = (b[raw[y][x + 1]] - b[raw[y][x - 1]]) / 2;
        else gx = b[z] - b[raw[y][x - 1]];
        e = grad_eval(r->grad_kind, gx, gy);
    }
    if (r->bias) b_add = r->bias[z] / r->w_start;
    r->en[z] = e + b_add;
}

static LqrRetVal build_emap(LqrCarver *r)
{
    gint x, y;
    if (r->nrg_uptodate) return LQR_OK;
    if (!r->rcache) LQR_CATCH(build_rcache(r));
    for (y = 0; y < r->h; y++)
        for (x = 0; x < r->w; x++) compute_e(r, x, y);
    r->nrg_uptodate = TRUE;
    return LQR_OK;
}

/* A.8 energy band after a carve; vpath_x is in pre-carve coordinates, w already decremented */
static LqrRetVal update_emap(LqrCarver *r)
{
    gint x, y, y1;
    gint rad = r->nrg_radius;
    if (r->nrg_uptodate) return LQR_OK;
    LQR_CATCH_F(r->rcache != NULL);
    for (y = 0; y < r->h; y++) {
        x = r->vpath_x[y];
        r->nrg_xmin[y] = x;
        r->nrg_xmax[y] = x - 1;
    }
    for (y = 0; y < r->h; y++) {
        gint y1_min = OMAX(y - rad, 0), y1_max = OMIN(y + rad, r->h - 1);
        x = r->vpath_x[y];
        for (y1 = y1_min; y1 <= y1_max; y1++) {
            r->nrg_xmin[y1] = OMAX(0, OMIN(r->nrg_xmin[y1], x - rad));
            r->nrg_xmax[y1] = OMIN(r->w - 1, OMAX(r->nrg_xmax[y1], x + rad - 1));
        }
    }
    for (y = 0; y < r->h; y++)
        for (x = r->nrg_xmin[y]; x <= r->nrg_xmax[y]; x++) compute_e(r, x, y);
    r->nrg_uptodate = TRUE;
    return LQR_OK;
}

/* ------------------------------------------------------------------ m-map DP (A.5) */
/* best parent of cell (x,y) among x+[-delta_x,delta_x] clipped; returns candidate value and parent id */
static inline gfloat best_parent(const LqrCarver *r, gint x, gint y, gint z, gint *parent)
{
    gint x1_min = OMAX(-x, -r->delta_x);
    gint x1_max = OMIN(r->w - 1 - x, r->delta_x);
    const gint *up = r->raw[y - 1];
    gint x1, zd = up[x + x1_min], least = zd;
    gfloat best, cand;
    if (r->rigidity) {
        gfloat r_fact = r->rigmask ? r->rigmask[z] : 1;
        best = r->m[zd] + r_fact * r->rigmap[x1_min];
        for (x1 = x1_min + 1; x1 <= x1_max; x1++) {
            zd = up[x + x1];
            cand = r->m[zd] + r_fact * r->rigmap[x1];
            if (cand < best || (cand == best && r->leftright == 1)) {
                best = cand;
                least = zd;
            }
        }
    } else {
        best = r->m[zd];
        for (x1 = x1_min + 1; x1 <= x1_max; x1++) {
            zd = up[x + x1];
            cand = r->m[zd];
            if (cand < best || (cand == best && r->leftright == 1)) {
                best = cand;
                least = zd;
            }
        }
    }
    *parent = least;
    return best;
}

static LqrRetVal build_mmap(LqrCarver *r)
{
    gint x, y;
    for (x = 0; x < r->w; x++) {
        gint z = r->raw[0][x];
        r->m[z] = r->en[z];
    }
    for (y = 1; y < r->h; y++)
        for (x = 0; x < r->w; x++) {
            gint z = r->raw[y][x], parent;
            gfloat best = best_parent(r, x, y, z, &parent);
            r->least[z] = parent;
            r->m[z] = r->en[z] + best;
        }
    return LQR_OK;
}

/* A.8 incremental DP with the keep-old rule and the self-trimming band */
static LqrRetVal update_mmap(LqrCarver *r)
{
    gint x, y, x_min, x_max;
    x_min = OMAX(r->nrg_xmin[0], 0);
    x_max = OMIN(r->nrg_xmax[0], r->w - 1);
    for (x = x_min; x <= x_max; x++) {
        gint z = r->raw[0][x];
        r->m[z] = r->en[z];
    }
    for (y = 1; y < r->h; y++) {
        gint stop = 0, x_stop = 0;
        x_min = OMIN(x_min, r->nrg_xmin[y]);
        x_max = OMAX(x_max, r->nrg_xmax[y]);
        x_min = OMAX(x_min - r->delta_x, 0);
        x_max = OMIN(x_max + r->delta_x, r->w - 1);
        r->stat_update_rows++;
        if (x_max >= x_min) {
            r->stat_update_cells += x_max - x_min + 1;
            if (x_max - x_min + 1 > r->stat_update_maxband) r->stat_update_maxband = x_max - x_min + 1;
            {
                gint b = 0, bw = x_max - x_min + 1;
                while ((1 << b) < bw && b < 15) b++;
                r->stat_band_hist[b]++;
            }
        }
        for (x = x_min; x <= x_max; x++) {
            gint z = r->raw[y][x], parent;
            gfloat new_m = r->en[z] + best_parent(r, x, y, z, &parent);
            if (r->least[z] == parent) {
                if (fabsf(r->m[z] - new_m) < UPDATE_TOLERANCE) {
                    if (!stop) x_stop = x;
                    stop = 1; /* m[z] kept */
                } else {
                    stop = 0;
                    r->m[z] = new_m;
                }
                if (x == x_min && stop) x_min++;
            } else {
                stop = 0;
                r->m[z] = new_m;
            }
            r->least[z] = parent;
            if (x == x_max && stop) x_max = x_stop;
        }
    }
    return LQR_OK;
}

/* ------------------------------------------------------------------ seam (A.6) */
static void build_vpath(LqrCarver *r)
{
    gint x, y = r->h - 1, last = -1, last_x = 0;
    gfloat best = (gfloat) (1 << 29);
    for (x = 0; x < r->w; x++) {
        gfloat v = r->m[r->raw[y][x]];
        if (v < best || (v == best && r->leftright == 1)) {
            last = r->raw[y][x];
            last_x = x;
            best = v;
        }
    }
    if (last < 0) { /* every m >= 2^29 or NaN: not reachable with 8-bit inputs and finite bias */
        last = r->raw[y][0];
        last_x = 0;
    }
    for (y = r->h0 - 1; y >= 0; y--) {
        r->vpath[y] = last;
        r->vpath_x[y] = last_x;
        if (y > 0) {
            gint x_min = OMAX(last_x - r->delta_x, 0), x_max = OMIN(last_x + r->delta_x, r->w - 1);
            last = r->least[r->raw[y][last_x]];
            for (x = x_min; x <= x_max; x++)
                if (r->raw[y - 1][x] == last) {
                    last_x = x;
                    break;
                }
        }
    }
}

static void update_vsmap(LqrCarver *r, gint l)
{
    gint y;
    for (y = 0; y < r->h; y++) r->vs[r->vpath[y]] = l;
}

/* A.7: index-table shift; w was already decremented */
static void carve(LqrCarver *r)
{
    gint x, y;
    for (y = 0; y < r->h_start; y++) {
        gint *row = r->raw[y];
        for (x = r->vpath_x[y]; x < r->w; x++) row[x] = row[x + 1];
    }
    r->nrg_uptodate = FALSE;
}

static void finish_vsmap(LqrCarver *r)
{
    gint y;
    cursor_reset(r);
    for (y = 1; y <= r->h; y++, cursor_next(r)) r->vs[r->cur_now] = r->w0;
    cursor_reset(r);
}

/* ------------------------------------------------------------------ inflate (A.9) */
static LqrRetVal inflate(LqrCarver *r, gint l)
{
    gint w1, z0, vs, k, x = 0, y = 0, c = r->channels;
    guchar *new_rgb;
    gint *new_vs = NULL;
    gfloat *new_bias = NULL, *new_rigmask = NULL;
    size_t n1;

    LQR_CATCH(foreach_attached(r, inflate, l));

    set_width(r, r->w0);
    w1 = r->w0 + l - r->max_level + 1;
    n1 = (size_t) w1 * r->h0;

    LQR_CATCH_MEM(new_rgb = (guchar *) calloc(n1 * c, 1));
    if (!r->root) LQR_CATCH_MEM(new_vs = (gint *) calloc(n1, sizeof(gint)));
    if (r->active) {
        if (r->bias) LQR_CATCH_MEM(new_bias = (gfloat *) calloc(n1, sizeof(gfloat)));
        if (r->rigmask) LQR_CATCH_MEM(new_rigmask = (gfloat *) calloc(n1, sizeof(gfloat)));
    }

    cursor_reset(r);
    for (z0 = 0; z0 < w1 * r->h0; z0++, cursor_next(r)) {
        gint now = r->cur_now;
        vs = r->vs[now];
        if (vs != 0 && vs <= l + r->max_level - 1 && vs >= 2 * r->max_level - 1) {
            /* a seam found in this session: insert a duplicate = integer mean with the left neighbour */
            gint left = r->cur_x > 0 ? cursor_left(r) : now;
            for (k = 0; k < c; k++) {
                gdouble t = (r->rgb[left * c + k] + r->rgb[now * c + k]) / 2; /* integer division, then widened */
                new_rgb[z0 * c + k] = (guchar) (t + 0.499999);
            }
            if (new_bias) new_bias[z0] = (r->bias[left] + r->bias[now]) / 2;
            if (new_rigmask) new_rigmask[z0] = (r->rigmask[left] + r->rigmask[now]) / 2;
            if (!r->root) new_vs[z0] = l - vs + r->max_level;
            z0++;
        }
        for (k = 0; k < c; k++) new_rgb[z0 * c + k] = r->rgb[now * c + k];
        if (new_bias) new_bias[z0] = r->bias[now];
        if (new_rigmask) new_rigmask[z0] = r->rigmask[now];
        if (vs != 0) {
            if (!r->root) new_vs[z0] = vs + l - r->max_level + 1;
        } else if (r->raw) {
            r->raw[y][x] = z0;
            x++;
            if (x >= r->w_start - l) {
                x = 0;
                y++;
            }
        }
    }

    free(r->rgb);
    free(r->en);
    free(r->m);
    free(r->rcache);
    free(r->least);
    free(r->bias);
    free(r->rigmask);
    r->en = r->m = r->bias = r->rigmask = NULL;
    r->least = NULL;
    r->rcache = NULL;
    r->nrg_uptodate = FALSE;
    r->rgb = new_rgb;
    if (!r->root) {
        free(r->vs);
        r->vs = new_vs;
        propagate_vs(r);
    }
    if (r->nrg_active) LQR_CATCH_MEM(r->en = (gfloat *) calloc(n1, sizeof(gfloat)));
    if (r->active) {
        r->bias = new_bias;
        r->rigmask = new_rigmask;
        LQR_CATCH_MEM(r->m = (gfloat *) calloc(n1, sizeof(gfloat)));
        LQR_CATCH_MEM(r->least = (gint *) calloc(n1, sizeof(gint)));
    }
    r->w0 = w1;
    r->w = r->w_start;
    r->level = l + 1;
    r->max_level = l + 1;
    free(r->line);
    LQR_CATCH_MEM(r->line = (guchar *) calloc((size_t) r->w0 * c, 1));
    cursor_reset(r);
    return LQR_OK;
}

/* ------------------------------------------------------------------ per-seam loop (A.7) */
static LqrRetVal build_vsmap(LqrCarver *r, gint depth)
{
    gint l, lr_switch_interval = 0;
    if (depth == 0) depth = r->w_start + 1;
    if (r->lr_switch_frequency) lr_switch_interval = (depth - r->max_level - 1) / (gint) r->lr_switch_frequency + 1;

    for (l = r->max_level; l < depth; l++) {
        gint done = l - r->max_level + r->session_rescale_current;
        if (done % r->session_update_step == 0)
            progress_update(r->progress, (gdouble) done / (gdouble) r->session_rescale_total);

        build_vpath(r);
        update_vsmap(r, l + r->max_level - 1);
        r->level++;
        r->w--;
        carve(r);

        if (r->w > 1) {
            LQR_CATCH(update_emap(r));
            if (r->lr_switch_frequency && ((l - r->max_level + lr_switch_interval / 2) % lr_switch_interval) == 0) {
                r->leftright ^= 1;
                LQR_CATCH(build_mmap(r));
            } else {
                LQR_CATCH(update_mmap(r));
            }
        } else {
            finish_vsmap(r);
        }
    }

    LQR_CATCH(inflate(r, depth - 1));
    set_width(r, r->w_start);
    LQR_CATCH(foreach_attached(r, set_width_attached, r->w_start));
    return LQR_OK;
}

static LqrRetVal build_maps(LqrCarver *r, gint depth)
{
    if (depth > r->max_level) {
        LQR_CATCH_F(r->active);
        LQR_CATCH_F(r->root == NULL);
        set_width(r, r->w_start - r->max_level + 1);
        LQR_CATCH(build_emap(r));
        LQR_CATCH(build_mmap(r));
        LQR_CATCH(build_vsmap(r, depth));
    }
    return LQR_OK;
}

/* ------------------------------------------------------------------ flatten / transpose (A.11) */
static LqrRetVal flatten_one(LqrCarver *r, gint unused)
{
    gint x, y, k, c = r->channels;
    guchar *new_rgb;
    gfloat *new_bias = NULL, *new_rigmask = NULL;
    size_t n = (size_t) r->w * r->h;
    (void) unused;

    LQR_CATCH(foreach_attached(r, flatten_one, 0));

    free(r->en);
    free(r->m);
    free(r->rcache);
    free(r->least);
    r->en = r->m = NULL;
    r->least = NULL;
    r->rcache = NULL;
    r->nrg_uptodate = FALSE;

    LQR_CATCH_MEM(new_rgb = (guchar *) calloc(n * c, 1));
    if (r->active && r->rigmask) LQR_CATCH_MEM(new_rigmask = (gfloat *) calloc(n, sizeof(gfloat)));
    if (r->nrg_active) {
        if (r->bias) LQR_CATCH_MEM(new_bias = (gfloat *) calloc(n, sizeof(gfloat)));
        free(r->raw_store);
        free(r->raw);
        LQR_CATCH_MEM(r->raw_store = (gint *) malloc(n * sizeof(gint)));
        LQR_CATCH_MEM(r->raw = (gint **) malloc(r->h * sizeof(gint *)));
    }

    cursor_reset(r);
    for (y = 0; y < r->h; y++) {
        if (r->nrg_active) r->raw[y] = r->raw_store + (size_t) y * r->w;
        for (x = 0; x < r->w; x++) {
            gint z0 = y * r->w + x, now = r->cur_now;
            for (k = 0; k < c; k++) new_rgb[z0 * c + k] = r->rgb[now * c + k];
            if (new_rigmask) new_rigmask[z0] = r->rigmask[now];
            if (r->nrg_active) {
                if (new_bias) new_bias[z0] = r->bias[now];
                r->raw[y][x] = z0;
            }
            cursor_next(r);
        }
    }

    free(r->rgb);
    r->rgb = new_rgb;
    if (r->nrg_active) {
        free(r->bias);
        r->bias = new_bias;
    }
    if (r->active) {
        free(r->rigmask);
        r->rigmask = new_rigmask;
    }
    if (!r->root) {
        free(r->vs);
        LQR_CATCH_MEM(r->vs = (gint *) calloc(n, sizeof(gint)));
        propagate_vs(r);
    }
    if (r->nrg_active) LQR_CATCH_MEM(r->en = (gfloat *) calloc(n, sizeof(gfloat)));
    if (r->active) {
        LQR_CATCH_MEM(r->m = (gfloat *) calloc(n, sizeof(gfloat)));
        LQR_CATCH_MEM(r->least = (gint *) calloc(n, sizeof(gint)));
    }
    r->w0 = r->w;
    r->h0 = r->h;
    r->w_start = r->w;
    r->h_start = r->h;
    r->level = 1;
    r->max_level = 1;
    free(r->line);
    LQR_CATCH_MEM(r->line = (guchar *) calloc((size_t) r->w0 * c, 1));
    cursor_reset(r);
    return LQR_OK;
}

static LqrRetVal transpose_one(LqrCarver *r, gint unused)
{
    gint x, y, k, d, c = r->channels;
    guchar *new_rgb;
    gfloat *new_bias = NULL, *new_rigmask = NULL;
    size_t n;
    (void) unused;

    if (r->level > 1) LQR_CATCH(flatten_one(r, 0));
    LQR_CATCH(foreach_attached(r, transpose_one, 0));

    n = (size_t) r->w0 * r->h0;
    if (!r->root) free(r->vs);
    free(r->en);
    free(r->m);
    free(r->rcache);
    free(r->least);
    free(r->line);
    r->en = r->m = NULL;
    r->least = NULL;
    r->rcache = NULL;
    r->line = NULL;
    r->nrg_uptodate = FALSE;

    LQR_CATCH_MEM(new_rgb = (guchar *) calloc(n * c, 1));
    if (!r->root) {
        LQR_CATCH_MEM(r->vs = (gint *) calloc(n, sizeof(gint)));
        propagate_vs(r);
    }
    if (r->nrg_active) {
        LQR_CATCH_MEM(r->en = (gfloat *) calloc(n, sizeof(gfloat)));
        if (r->bias) LQR_CATCH_MEM(new_bias = (gfloat *) calloc(n, sizeof(gfloat)));
        free(r->raw_store);
        free(r->raw);
        LQR_CATCH_MEM(r->raw_store = (gint *) calloc(n, sizeof(gint)));
        LQR_CATCH_MEM(r->raw = (gint **) calloc(r->w0, sizeof(gint *)));
    }
    if (r->active) {
        LQR_CATCH_MEM(r->m = (gfloat *) calloc(n, sizeof(gfloat)));
        LQR_CATCH_MEM(r->least = (gint *) calloc(n, sizeof(gint)));
        if (r->rigmask) LQR_CATCH_MEM(new_rigmask = (gfloat *) calloc(n, sizeof(gfloat)));
    }

    for (x = 0; x < r->w; x++) {
        if (r->nrg_active) r->raw[x] = r->raw_store + (size_t) x * r->h0;
        for (y = 0; y < r->h; y++) {
            gint z0 = y * r->w0 + x, z1 = x * r->h0 + y;
            for (k = 0; k < c; k++) new_rgb[z1 * c + k] = r->rgb[z0 * c + k];
            if (new_bias) new_bias[z1] = r->bias[z0];
            if (new_rigmask) new_rigmask[z1] = r->rigmask[z0];
            if (r->nrg_active) r->raw[x][y] = z1;
        }
    }

    free(r->rgb);
    r->rgb = new_rgb;
    if (r->nrg_active) {
        free(r->bias);
        r->bias = new_bias;
    }
    if (r->active) {
        free(r->rigmask);
        r->rigmask = new_rigmask;
    }

    d = r->w0;
    r->w0 = r->h0;
    r->h0 = d;
    r->w = r->w0;
    r->h = r->h0;
    r->w_start = r->w0;
    r->h_start = r->h0;
    r->level = 1;
    r->max_level = 1;

    if (r->active) {
        free(r->vpath);
        free(r->vpath_x);
        free(r->nrg_xmin);
        free(r->nrg_xmax);
        LQR_CATCH_MEM(r->vpath = (gint *) malloc(r->h * sizeof(gint)));
        LQR_CATCH_MEM(r->vpath_x = (gint *) malloc(r->h * sizeof(gint)));
        LQR_CATCH_MEM(r->nrg_xmin = (gint *) malloc(r->h * sizeof(gint)));
        LQR_CATCH_MEM(r->nrg_xmax = (gint *) malloc(r->h * sizeof(gint)));
    }
    LQR_CATCH_MEM(r->line = (guchar *) calloc((size_t) r->w0 * c, 1));

    if (r->active)
        for (x = -r->delta_x; x <= r->delta_x; x++) r->rigmap[x] = r->rigmap[x] * r->w0 / r->h0;

    r->transposed = r->transposed ? 0 : 1;
    cursor_reset(r);
    return LQR_OK;
}

/* ------------------------------------------------------------------ vmaps (A.13) */
static LqrVMap *vmap_snapshot(LqrCarver *r)
{
    gint w1 = r->w, w, h, x, y, depth;
    gint *buffer;
    LqrVMap *vmap;

    set_width(r, r->w_start);
    w = lqr_carver_get_width(r);
    h = lqr_carver_get_height(r);
    depth = r->w0 - r->w_start;
    buffer = (gint *) malloc(sizeof(gint) * (size_t) w * h);
    if (!buffer) return NULL;

    cursor_reset(r);
    for (y = 0; y < r->h; y++)
        for (x = 0; x < r->w; x++) {
            gint vs = r->vs[r->cur_now];
            gint z0 = r->transposed ? x * r->h + y : y * r->w + x;
            buffer[z0] = vs == 0 ? 0 : vs - depth;
            cursor_next(r);
        }
    set_width(r, w1);
    cursor_reset(r);

    vmap = (LqrVMap *) malloc(sizeof(LqrVMap));
    if (!vmap) {
        free(buffer);
        return NULL;
    }
    vmap->buffer = buffer;
    vmap->width = w;
    vmap->height = h;
    vmap->depth = depth;
    vmap->orientation = r->transposed;
    return vmap;
}

LqrVMap *lqr_vmap_dump(LqrCarver *r) { return r ? vmap_snapshot(r) : NULL; }

static LqrRetVal vmap_internal_dump(LqrCarver *r)
{
    LqrVMap *vmap = vmap_snapshot(r);
    LqrVMapList *node, **tail;
    LQR_CATCH_MEM(vmap);
    node = (LqrVMapList *) malloc(sizeof(LqrVMapList));
    LQR_CATCH_MEM(node);
    node->current = vmap;
    node->next = NULL;
    for (tail = &r->flushed_vs; *tail; tail = &(*tail)->next) {}
    *tail = node;
    return LQR_OK;
}

void lqr_vmap_destroy(LqrVMap *v)
{
    if (!v) return;
    free(v->buffer);
    free(v);
}
gint *lqr_vmap_get_data(LqrVMap *v) { return v->buffer; }
gint lqr_vmap_get_width(LqrVMap *v) { return v->width; }
gint lqr_vmap_get_height(LqrVMap *v) { return v->height; }
gint lqr_vmap_get_depth(LqrVMap *v) { return v->depth; }
gint lqr_vmap_get_orientation(LqrVMap *v) { return v->orientation; }
LqrVMapList *lqr_vmap_list_start(LqrCarver *r) { return r->flushed_vs; }
LqrVMap *lqr_vmap_list_current(LqrVMapList *l) { return l ? l->current : NULL; }
LqrVMapList *lqr_vmap_list_next(LqrVMapList *l) { return l ? l->next : NULL; }
LqrRetVal lqr_vmap_list_foreach(LqrVMapList *l, LqrVMapFunc func, gpointer data)
{
    for (; l; l = l->next) LQR_CATCH(func(l->current, data));
    return LQR_OK;
}

/* ------------------------------------------------------------------ public: life cycle */
LqrCarver *lqr_carver_new(guchar *buffer, gint width, gint height, gint channels)
{
    LqrCarver *r;
    if (!buffer || width < 1 || height < 1 || channels < 1 || channels > 4) return NULL;
    r = (LqrCarver *) calloc(1, sizeof(LqrCarver));
    if (!r) return NULL;
    r->level = r->max_level = 1;
    r->resize_order = LQR_RES_ORDER_HOR;
    r->progress = lqr_progress_new();
    r->session_update_step = 1;
    r->delta_x = 1;
    r->w = r->w0 = r->w_start = width;
    r->h = r->h0 = r->h_start = height;
    r->channels = channels;
    r->alpha_channel = (channels == 2 || channels == 4) ? channels - 1 : -1;
    r->rgb = buffer;
    lqr_carver_set_energy_function_builtin(r, LQR_EF_GRAD_XABS);
    r->enl_step = 2.0f;
    r->vs = (gint *) calloc((size_t) width * height, sizeof(gint));
    r->line = (guchar *) calloc((size_t) width * channels, 1);
    if (!r->progress || !r->vs || !r->line) {
        r->rgb = NULL; /* ownership not taken on failure */
        lqr_carver_destroy(r);
        return NULL;
    }
    cursor_reset(r);
    return r;
}

void lqr_carver_destroy(LqrCarver *r)
{
    LqrCarverList *it, *itn;
    LqrVMapList *vl, *vln;
    if (!r) return;
    for (it = r->attached; it; it = itn) {
        itn = it->next;
        lqr_carver_destroy(it->current);
        free(it);
    }
    for (vl = r->flushed_vs; vl; vl = vln) {
        vln = vl->next;
        lqr_vmap_destroy(vl->current);
        free(vl);
    }
    free(r->rgb);
    if (!r->root) free(r->vs);
    free(r->en);
    free(r->bias);
    free(r->m);
    free(r->rigmask);
    free(r->least);
    free(r->raw_store);
    free(r->raw);
    free(r->vpath);
    free(r->vpath_x);
    free(r->nrg_xmin);
    free(r->nrg_xmax);
    free(r->rcache);
    free(r->rigmap_store);
    free(r->line);
    free(r->progress);
    free(r);
}

static LqrRetVal init_energy_related(LqrCarver *r)
{
    gint x, y;
    size_t n = (size_t) r->w * r->h;
    LQR_CATCH_F(!r->active && !r->nrg_active);
    LQR_CATCH_MEM(r->en = (gfloat *) calloc(n, sizeof(gfloat)));
    LQR_CATCH_MEM(r->raw_store = (gint *) malloc(sizeof(gint) * (size_t) r->h_start * r->w_start));
    LQR_CATCH_MEM(r->raw = (gint **) malloc(sizeof(gint *) * r->h_start));
    for (y = 0; y < r->h; y++) {
        r->raw[y] = r->raw_store + (size_t) y * r->w_start;
        for (x = 0; x < r->w_start; x++) r->raw[y][x] = y * r->w_start + x;
    }
    r->nrg_active = TRUE;
    return LQR_OK;
}

LqrRetVal lqr_carver_init(LqrCarver *r, gint delta_x, gfloat rigidity)
{
    gint x;
    size_t n;
    LQR_CATCH_F(r != NULL && delta_x >= 0);
    LQR_CATCH_F(r->active == FALSE);
    if (!r->nrg_active) LQR_CATCH(init_energy_related(r));
    n = (size_t) r->w * r->h;
    LQR_CATCH_MEM(r->m = (gfloat *) calloc(n, sizeof(gfloat)));
    LQR_CATCH_MEM(r->least = (gint *) calloc(n, sizeof(gint)));
    LQR_CATCH_MEM(r->vpath = (gint *) malloc(r->h * sizeof(gint)));
    LQR_CATCH_MEM(r->vpath_x = (gint *) malloc(r->h * sizeof(gint)));
    LQR_CATCH_MEM(r->nrg_xmin = (gint *) malloc(r->h * sizeof(gint)));
    LQR_CATCH_MEM(r->nrg_xmax = (gint *) malloc(r->h * sizeof(gint)));
    r->delta_x = delta_x;
    r->rigidity = rigidity;
    LQR_CATCH_MEM(r->rigmap_store = (gfloat *) calloc(2 * delta_x + 1, sizeof(gfloat)));
    r->rigmap = r->rigmap_store + delta_x;
    for (x = -delta_x; x <= delta_x; x++) r->rigmap[x] = r->rigidity * powf(fabsf((gfloat) x), 1.5f) / r->h;
    r->active = TRUE;
    return LQR_OK;
}

LqrRetVal lqr_carver_attach(LqrCarver *r, LqrCarver *aux)
{
    LqrCarverList *node, **tail;
    LQR_CATCH_F(r != NULL && aux != NULL);
    LQR_CATCH_F(r->w0 == aux->w0);
    LQR_CATCH_F(r->h0 == aux->h0);
    LQR_CATCH_MEM(node = (LqrCarverList *) malloc(sizeof(LqrCarverList)));
    node->current = aux;
    node->next = NULL;
    for (tail = &r->attached; *tail; tail = &(*tail)->next) {}
    *tail = node;
    free(aux->vs);
    aux->vs = r->vs;
    aux->root = r;
    return LQR_OK;
}

/* ------------------------------------------------------------------ public: knobs */
LqrRetVal lqr_carver_set_energy_function_builtin(LqrCarver *r, LqrEnergyFuncBuiltinType ef)
{
    gint grad, rd, rad = 1;
    LQR_CATCH_F(r != NULL);
    switch (ef) {
        case LQR_EF_GRAD_NORM: grad = GRAD_NORM; rd = READ_BRIGHTNESS; break;
        case LQR_EF_GRAD_SUMABS: grad = GRAD_SUMABS; rd = READ_BRIGHTNESS; break;
        case LQR_EF_GRAD_XABS: grad = GRAD_XABS; rd = READ_BRIGHTNESS; break;
        case LQR_EF_LUMA_GRAD_NORM: grad = GRAD_NORM; rd = READ_LUMA; break;
        case LQR_EF_LUMA_GRAD_SUMABS: grad = GRAD_SUMABS; rd = READ_LUMA; break;
        case LQR_EF_LUMA_GRAD_XABS: grad = GRAD_XABS; rd = READ_LUMA; break;
        case LQR_EF_NULL: grad = GRAD_NULL; rd = READ_BRIGHTNESS; rad = 0; break;
        default: return LQR_ERROR;
    }
    r->ef_index = ef;
    r->grad_kind = grad;
    r->read_kind = rd;
    r->nrg_radius = rad;
    free(r->rcache);
    r->rcache = NULL;
    r->nrg_uptodate = FALSE;
    return LQR_OK;
}
void lqr_carver_set_resize_order(LqrCarver *r, LqrResizeOrder o) { if (r) r->resize_order = o; }
void lqr_carver_set_progress(LqrCarver *r, LqrProgress *p)
{
    if (!r) return;
    free(r->progress);
    r->progress = p;
}
void lqr_carver_set_side_switch_frequency(LqrCarver *r, guint f) { if (r) r->lr_switch_frequency = f; }
LqrRetVal lqr_carver_set_enl_step(LqrCarver *r, gfloat s)
{
    LQR_CATCH_F(r != NULL);
    LQR_CATCH_F((s > 1) && (s <= 2));
    r->enl_step = s;
    return LQR_OK;
}
void lqr_carver_set_dump_vmaps(LqrCarver *r) { if (r) r->dump_vmaps = TRUE; }
void lqr_carver_set_no_dump_vmaps(LqrCarver *r) { if (r) r->dump_vmaps = FALSE; }

/* ------------------------------------------------------------------ public: getters */
gint lqr_carver_get_width(LqrCarver *r) { return r->transposed ? r->h : r->w; }
gint lqr_carver_get_height(LqrCarver *r) { return r->transposed ? r->w : r->h; }
gint lqr_carver_get_ref_width(LqrCarver *r) { return r->transposed ? r->h_start : r->w_start; }
gint lqr_carver_get_ref_height(LqrCarver *r) { return r->transposed ? r->w_start : r->h_start; }
gint lqr_carver_get_channels(LqrCarver *r) { return r->channels; }
gint lqr_carver_get_orientation(LqrCarver *r) { return r->transposed ? 1 : 0; }
gint lqr_carver_get_depth(LqrCarver *r) { return r->w0 - r->w_start; }
gfloat lqr_carver_get_enl_step(LqrCarver *r) { return r->enl_step; }

/* ------------------------------------------------------------------ public: masks (A.4) */
static int not_at_reference(const LqrCarver *r)
{
    return r->w != r->w0 || r->w_start != r->w0 || r->h != r->h0 || r->h_start != r->h0;
}

LqrRetVal lqr_carver_bias_add_rgb_area(LqrCarver *r, guchar *rgb, gint bias_factor, gint channels,
                                       gint width, gint height, gint x_off, gint y_off)
{
    gint x, y, k, c_channels, has_alpha, x0, y0, x1, y1, x2, y2, transposed;
    LQR_CATCH_F(r != NULL && rgb != NULL && channels >= 1);
    if (not_at_reference(r)) LQR_CATCH(flatten_one(r, 0));
    if (bias_factor == 0) return LQR_OK;
    if (!r->bias) LQR_CATCH_MEM(r->bias = (gfloat *) calloc((size_t) r->w * r->h, sizeof(gfloat)));
    has_alpha = (channels == 2 || channels >= 4);
    c_channels = channels - has_alpha;
    transposed = r->transposed;
    if (transposed) LQR_CATCH(transpose_one(r, 0));

    x0 = OMIN(0, x_off);
    y0 = OMIN(0, y_off);
    x1 = OMAX(0, x_off);
    y1 = OMAX(0, y_off);
    x2 = OMIN(r->w, width + x_off);
    y2 = OMIN(r->h, height + y_off);
    for (y = 0; y < y2 - y1; y++)
        for (x = 0; x < x2 - x1; x++) {
            gint sum = 0, px = (y - y0) * width + (x - x0);
            gfloat bias;
            for (k = 0; k < c_channels; k++) sum += rgb[px * channels + k];
            bias = (gfloat) ((gdouble) bias_factor * sum / (2 * 255 * c_channels));
            if (has_alpha) bias *= (gfloat) rgb[(px + 1) * channels - 1] / 255;
            r->bias[(y + y1) * r->w0 + (x + x1)] += bias;
        }
    r->nrg_uptodate = FALSE;
    if (transposed != r->transposed) LQR_CATCH(transpose_one(r, 0));
    return LQR_OK;
}

LqrRetVal lqr_carver_rigmask_add_rgb_area(LqrCarver *r, guchar *rgb, gint channels,
                                          gint width, gint height, gint x_off, gint y_off)
{
    gint x, y, k, c_channels, has_alpha, x0, y0, x1, y1, x2, y2, transposed;
    LQR_CATCH_F(r != NULL && rgb != NULL && channels >= 1);
    LQR_CATCH_F(r->active);
    if (not_at_reference(r)) LQR_CATCH(flatten_one(r, 0));
    if (!r->rigmask) LQR_CATCH_MEM(r->rigmask = (gfloat *) calloc((size_t) r->w0 * r->h0, sizeof(gfloat)));
    has_alpha = (channels == 2 || channels >= 4);
    c_channels = channels - has_alpha;
    transposed = r->transposed;
    if (transposed) LQR_CATCH(transpose_one(r, 0));

    x0 = OMIN(0, x_off);
    y0 = OMIN(0, y_off);
    x1 = OMAX(0, x_off);
    y1 = OMAX(0, y_off);
    x2 = OMIN(r->w, width + x_off);
    y2 = OMIN(r->h, height + y_off);
    for (y = 0; y < y2 - y1; y++)
        for (x = 0; x < x2 - x1; x++) {
            gint sum = 0, px = (y - y0) * width + (x - x0);
            gfloat v;
            for (k = 0; k < c_channels; k++) sum += rgb[px * channels + k];
            v = (gfloat) sum / (255 * c_channels);
            if (has_alpha) v *= (gfloat) rgb[(px + 1) * channels - 1] / 255;
            r->rigmask[(y + y1) * r->w0 + (x + x1)] = v;
        }
    if (transposed != r->transposed) LQR_CATCH(transpose_one(r, 0));
    return LQR_OK;
}

/* ------------------------------------------------------------------ public: resize driver (A.10) */
static gint step_limit(gfloat enl_step, gint ref)
{
    gint d = (gint) ((enl_step - 1) * ref) - 1;
    return d < 1 ? 1 : d;
}

/* one direction; `along_w` selects whether the request is for the image width (TRUE) or height */
static LqrRetVal resize_direction(LqrCarver *r, gint target, gboolean along_w)
{
    /* the direction is carved along internal x; the image must be transposed iff
     * (along_w && r->transposed) || (!along_w && !r->transposed) */
    gboolean need_flip = along_w ? r->transposed : !r->transposed;
    gint ref = need_flip ? r->h_start : r->w_start;
    gint cur = need_flip ? r->h : r->w;
    gint delta = target - ref, gamma = target - cur, delta_max = step_limit(r->enl_step, ref);
    const gchar *msg_init = along_w ? r->progress->init_width_message : r->progress->init_height_message;
    const gchar *msg_end = along_w ? r->progress->end_width_message : r->progress->end_height_message;

    if (delta < 0) {
        delta = -delta;
        delta_max = delta;
    }
    r->session_rescale_total = gamma > 0 ? gamma : -gamma;
    r->session_rescale_current = 0;
    r->session_update_step = (gint) OMAX(r->session_rescale_total * r->progress->update_step, 1);
    if (r->session_rescale_total) progress_init(r->progress, msg_init);

    while (gamma) {
        gint delta0 = OMIN(delta, delta_max), new_w;
        delta -= delta0;
        if (along_w ? r->transposed : !r->transposed) LQR_CATCH(transpose_one(r, 0));
        new_w = OMIN(target, r->w_start + delta_max);
        gamma = target - new_w;
        LQR_CATCH(build_maps(r, delta0 + 1));
        set_width(r, new_w);
        LQR_CATCH(foreach_attached(r, set_width_attached, new_w));
        r->session_rescale_current = r->session_rescale_total - (gamma > 0 ? gamma : -gamma);
        if (r->dump_vmaps) LQR_CATCH(vmap_internal_dump(r));
        if (new_w < target) {
            LQR_CATCH(flatten_one(r, 0));
            delta_max = step_limit(r->enl_step, r->w_start);
        }
    }
    if (r->session_rescale_total) progress_end(r->progress, msg_end);
    return LQR_OK;
}

static void scan_reset_all(LqrCarver *r)
{
    LqrCarverList *it;
    cursor_reset(r);
    for (it = r->attached; it; it = it->next) scan_reset_all(it->current);
}

LqrRetVal lqr_carver_resize(LqrCarver *r, gint w1, gint h1)
{
    LQR_CATCH_F(r != NULL);
    LQR_CATCH_F((w1 >= 1) && (h1 >= 1));
    LQR_CATCH_F(r->root == NULL);
    if (r->resize_order == LQR_RES_ORDER_HOR) {
        LQR_CATCH(resize_direction(r, w1, TRUE));
        LQR_CATCH(resize_direction(r, h1, FALSE));
    } else {
        LQR_CATCH(resize_direction(r, h1, FALSE));
        LQR_CATCH(resize_direction(r, w1, TRUE));
    }
    scan_reset_all(r);
    return LQR_OK;
}

LqrRetVal lqr_carver_flatten(LqrCarver *r)
{
    LQR_CATCH_F(r != NULL);
    return flatten_one(r, 0);
}

/* ------------------------------------------------------------------ public: read-out (A.12) */
void lqr_carver_scan_reset(LqrCarver *r) { if (r) cursor_reset(r); }

gboolean lqr_carver_scan_by_row(LqrCarver *r) { return r->transposed ? FALSE : TRUE; }

gboolean lqr_carver_scan_line(LqrCarver *r, gint *n, guchar **rgb)
{
    gint x, k, c = r->channels;
    if (r->cur_eoc) {
        cursor_reset(r);
        return FALSE;
    }
    /* the cursor always rests at the start of a line between calls */
    *n = r->cur_y;
    for (x = 0; x < r->w; x++) {
        for (k = 0; k < c; k++) r->line[x * c + k] = r->rgb[r->cur_now * c + k];
        cursor_next(r);
    }
    *rgb = r->line;
    return TRUE;
}

gboolean lqr_carver_scan(LqrCarver *r, gint *x, gint *y, guchar **rgb)
{
    gint k, c = r->channels;
    if (r->cur_eoc) {
        cursor_reset(r);
        return FALSE;
    }
    *x = r->transposed ? r->cur_y : r->cur_x;
    *y = r->transposed ? r->cur_x : r->cur_y;
    for (k = 0; k < c; k++) r->line[k] = r->rgb[r->cur_now * c + k];
    *rgb = r->line;
    cursor_next(r);
    return TRUE;
}

LqrRetVal lqr_carver_get_true_energy(LqrCarver *r, gfloat *buffer, gint orientation)
{
    gint x, y, w, h;
    LQR_CATCH_F(r != NULL && buffer != NULL);
    LQR_CATCH_F(orientation == 0 || orientation == 1);
    if (!r->nrg_active) LQR_CATCH(init_energy_related(r));
    if (r->w != r->w_start - r->max_level + 1) LQR_CATCH(flatten_one(r, 0));
    if (orientation != lqr_carver_get_orientation(r)) LQR_CATCH(transpose_one(r, 0));
    LQR_CATCH(build_emap(r));
    w = lqr_carver_get_width(r);
    h = lqr_carver_get_height(r);
    for (y = 0; y < h; y++)
        for (x = 0; x < w; x++) {
            gint z = orientation == 0 ? r->raw[y][x] : r->raw[x][y];
            buffer[y * w + x] = r->en[z];
        }
    return LQR_OK;
}

/* ------------------------------------------------------------------ public: lists */
LqrCarverList *lqr_carver_list_start(LqrCarver *r) { return r->attached; }
LqrCarver *lqr_carver_list_current(LqrCarverList *l) { return l ? l->current : NULL; }
LqrCarverList *lqr_carver_list_next(LqrCarverList *l) { return l ? l->next : NULL; }

/* ------------------------------------------------------------------ oracle-only instrumentation */
/* band statistics of the incremental DP (rows visited, cells recomputed, widest band) */
LQR_PUBLIC void lqr_oracle_update_stats(LqrCarver *r, long out[3])
{
    out[0] = r->stat_update_rows;
    out[1] = r->stat_update_cells;
    out[2] = r->stat_update_maxband;
}

LQR_PUBLIC void lqr_oracle_band_hist(LqrCarver *r, long out[16])
{
    memcpy(out, r->stat_band_hist, sizeof r->stat_band_hist);
}

/* state after `n_seams` iterations of the per-seam loop WITHOUT the final inflate, so that the maps can
 * be compared with the CUDA engine's (mirrors b200c_debug_build); the carver is not resizable afterwards */
LQR_PUBLIC LqrRetVal lqr_oracle_debug_build(LqrCarver *r, gint n_seams)
{
    gint l, depth, lr_switch_interval = 0, first;
    LQR_CATCH_F(r != NULL && r->active && r->root == NULL);
    set_width(r, r->w_start - r->max_level + 1);
    LQR_CATCH(build_emap(r));
    LQR_CATCH(build_mmap(r));
    first = r->max_level;
    depth = first + n_seams;
    if (r->lr_switch_frequency) lr_switch_interval = (depth - r->max_level - 1) / (gint) r->lr_switch_frequency + 1;
    for (l = first; l < depth; l++) {
        build_vpath(r);
        update_vsmap(r, l + r->max_level - 1);
        r->level++;
        r->w--;
        carve(r);
        if (r->w > 1) {
            LQR_CATCH(update_emap(r));
            if (r->lr_switch_frequency && ((l - r->max_level + lr_switch_interval / 2) % lr_switch_interval) == 0) {
                r->leftright ^= 1;
                LQR_CATCH(build_mmap(r));
            } else {
                LQR_CATCH(update_mmap(r));
            }
        } else {
            finish_vsmap(r);
        }
    }
    return LQR_OK;
}

/* what: 0 en, 1 m, 2 least, 3 raw (w_start*h_start, row pitch w_start), 4 vs, 5 vpath_x, 6 bias, 7 rigmask */
LQR_PUBLIC long lqr_oracle_debug_fetch(LqrCarver *r, gint what, void *out, long cap)
{
    const void *src = NULL;
    long n = (long) r->w0 * r->h0;
    switch (what) {
        case 0: src = r->en; break;
        case 1: src = r->m; break;
        case 2: src = r->least; break;
        case 3: src = r->raw_store; n = (long) r->w_start * r->h_start; break;
        case 4: src = r->vs; break;
        case 5: src = r->vpath_x; n = r->h; break;
        case 6: src = r->bias; break;
        case 7: src = r->rigmask; break;
        default: return -1;
    }
    if (!src) return 0;
    if (n > cap) n = cap;
    memcpy(out, src, (size_t) n * 4);
    return n;
}
