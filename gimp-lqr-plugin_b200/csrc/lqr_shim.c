/* lqr_shim.c -- liblqr-1.so: the plain-C drop-in the plug-in links against (LDADD = $(LQR_LIBS),
 * reference src/Makefile.am:36,39).  It exports include/lqr.h and owns everything that is host
 * business in the LqrCarver API: handle bookkeeping, the resize driver (order, enlargement stepping,
 * SURVEY.md A.10), progress callbacks on the caller's thread (render.c:767-779), attached-carver and
 * seam-map lists, buffer ownership (render.c:220-223) and the line cursor of lqr_carver_scan_line
 * (io_functions.c:155-164).  All pixel arithmetic is delegated to the CUDA engine libb200carve.so
 * (include/b200carve.h), which is dlopen()ed on first use.  There is no CPU fallback: if the engine or a
 * CUDA device is missing, lqr_carver_new() returns NULL and says why on stderr.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "b200carve.h"
#include "lqr.h"

/* ------------------------------------------------------------------ engine binding */
typedef struct {
    void *dl;
    int (*abi_version)(void);
    const char *(*last_error)(void);
    B200Carver *(*carver_new)(const unsigned char *, int, int, int);
    void (*carver_destroy)(B200Carver *);
    int (*carver_init)(B200Carver *, int, float);
    int (*carver_attach)(B200Carver *, B200Carver *);
    int (*set_energy_function)(B200Carver *, int);
    int (*set_side_switch_frequency)(B200Carver *, unsigned int);
    int (*bias_add_rgb_area)(B200Carver *, const unsigned char *, int, int, int, int, int, int);
    int (*rigmask_add_rgb_area)(B200Carver *, const unsigned char *, int, int, int, int, int);
    int (*build_maps)(B200Carver *, int, int, b200c_progress_fn, void *);
    int (*build_maps_phase)(B200Carver *, int, int, int, b200c_progress_fn, void *);
    int (*batch_build_maps)(B200Carver **, int, int);
    int (*set_width)(B200Carver *, int);
    int (*flatten)(B200Carver *);
    int (*transpose)(B200Carver *);
    int (*get)(const B200Carver *, int);
    int (*readout)(B200Carver *, const unsigned char **);
    int (*readout_rows)(B200Carver *, int);
    int (*vmap)(B200Carver *, int *);
    int (*true_energy)(B200Carver *, float *);
} Engine;

static Engine g_eng;
static int g_eng_state = 0; /* 0 = not tried, 1 = ok, -1 = failed */

static void *try_open(const char *path)
{
    return path ? dlopen(path, RTLD_NOW | RTLD_LOCAL) : NULL;
}

#define BIND(field, sym)                                                          \
    do {                                                                          \
        *(void **) (&g_eng.field) = dlsym(g_eng.dl, sym);                         \
        if (!g_eng.field) {                                                       \
            fprintf(stderr, "liblqr-1 (b200): engine lacks symbol %s\n", sym);    \
            return 0;                                                             \
        }                                                                         \
    } while (0)

static int engine_load_once(void)
{
    Dl_info info;
    char path[4096];
    g_eng.dl = try_open(getenv("B200CARVE_LIB"));
    if (!g_eng.dl && dladdr((void *) &engine_load_once, &info) && info.dli_fname) {
        const char *slash = strrchr(info.dli_fname, '/');
        size_t dir = slash ? (size_t) (slash - info.dli_fname) + 1 : 0;
        if (dir + 32 < sizeof path) {
            memcpy(path, info.dli_fname, dir);
            strcpy(path + dir, "libb200carve.so");
            g_eng.dl = try_open(path);
        }
    }
    if (!g_eng.dl) g_eng.dl = try_open("libb200carve.so");
    if (!g_eng.dl) {
        fprintf(stderr, "liblqr-1 (b200): cannot load the CUDA engine libb200carve.so: %s\n", dlerror());
        return 0;
    }
    BIND(abi_version, "b200c_abi_version");
    BIND(last_error, "b200c_last_error");
    BIND(carver_new, "b200c_carver_new");
    BIND(carver_destroy, "b200c_carver_destroy");
    BIND(carver_init, "b200c_carver_init");
    BIND(carver_attach, "b200c_carver_attach");
    BIND(set_energy_function, "b200c_carver_set_energy_function");
    BIND(set_side_switch_frequency, "b200c_carver_set_side_switch_frequency");
    BIND(bias_add_rgb_area, "b200c_carver_bias_add_rgb_area");
    BIND(rigmask_add_rgb_area, "b200c_carver_rigmask_add_rgb_area");
    BIND(build_maps, "b200c_carver_build_maps");
    BIND(build_maps_phase, "b200c_carver_build_maps_phase");
    BIND(batch_build_maps, "b200c_batch_build_maps");
    BIND(set_width, "b200c_carver_set_width");
    BIND(flatten, "b200c_carver_flatten");
    BIND(transpose, "b200c_carver_transpose");
    BIND(get, "b200c_carver_get");
    BIND(readout, "b200c_carver_readout_begin");
    BIND(readout_rows, "b200c_carver_readout_rows");
    BIND(vmap, "b200c_carver_vmap");
    BIND(true_energy, "b200c_carver_true_energy");
    if (g_eng.abi_version() != B200C_ABI_VERSION) {
        fprintf(stderr, "liblqr-1 (b200): engine ABI %d, shim built for %d\n", g_eng.abi_version(), B200C_ABI_VERSION);
        return 0;
    }
    return 1;
}

/* independent carvers may be driven from different host threads (a batch host): load exactly once */
static pthread_once_t g_eng_once = PTHREAD_ONCE_INIT;
static void engine_load_thunk(void) { g_eng_state = engine_load_once() ? 1 : -1; }
static int engine_load(void)
{
    pthread_once(&g_eng_once, engine_load_thunk);
    return g_eng_state > 0;
}

/* ------------------------------------------------------------------ handle types */
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

struct _LqrCarver {
    B200Carver *eng;
    LqrCarver *root;
    LqrCarverList *attached;
    LqrVMapList *flushed_vs;
    LqrProgress *progress;
    LqrResizeOrder resize_order;
    gboolean dump_vmaps;
    gfloat enl_step;
    gint channels;
    gint session_update_step, session_rescale_total, session_rescale_current;
    /* read-out: lines served from one device gather (b200c_carver_readout) */
    const guchar *lines; /* engine-owned pinned buffer, NULL when stale */
    gint line_w, line_h;  /* internal geometry of `lines` */
    gint cur_line, cur_x;
    gint lines_ready;     /* rows of `lines` that have arrived (the read-out comes back in chunks) */
    guchar pixel[4];
};

#define EG(r, f) (g_eng.get((r)->eng, (f)))

static void invalidate_lines(LqrCarver *r)
{
    LqrCarverList *it;
    r->lines = NULL;
    r->cur_line = 0;
    r->cur_x = 0;
    for (it = r->attached; it; it = it->next) invalidate_lines(it->current);
}

/* ------------------------------------------------------------------ progress */
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

LqrRetVal lqr_progress_set_init(LqrProgress *p, LqrProgressFuncInit f)
{
    LQR_CATCH_F(p != NULL);
    p->init = f;
    return LQR_OK;
}

LqrRetVal lqr_progress_set_update(LqrProgress *p, LqrProgressFuncUpdate f)
{
    LQR_CATCH_F(p != NULL);
    p->update = f;
    return LQR_OK;
}

LqrRetVal lqr_progress_set_end(LqrProgress *p, LqrProgressFuncEnd f)
{
    LQR_CATCH_F(p != NULL);
    p->end = f;
    return LQR_OK;
}

LqrRetVal lqr_progress_set_update_step(LqrProgress *p, gfloat s)
{
    LQR_CATCH_F(p != NULL);
    p->update_step = s;
    return LQR_OK;
}

static LqrRetVal copy_message(gchar *dst, const gchar *src)
{
    LQR_CATCH_F(src != NULL);
    snprintf(dst, LQR_PROGRESS_MAX_MESSAGE_LENGTH, "%s", src);
    return LQR_OK;
}

LqrRetVal lqr_progress_set_init_width_message(LqrProgress *p, const gchar *m)
{
    LQR_CATCH_F(p != NULL);
    return copy_message(p->init_width_message, m);
}

LqrRetVal lqr_progress_set_init_height_message(LqrProgress *p, const gchar *m)
{
    LQR_CATCH_F(p != NULL);
    return copy_message(p->init_height_message, m);
}

LqrRetVal lqr_progress_set_end_width_message(LqrProgress *p, const gchar *m)
{
    LQR_CATCH_F(p != NULL);
    return copy_message(p->end_width_message, m);
}

LqrRetVal lqr_progress_set_end_height_message(LqrProgress *p, const gchar *m)
{
    LQR_CATCH_F(p != NULL);
    return copy_message(p->end_height_message, m);
}

/* ------------------------------------------------------------------ life cycle */
LqrCarver *lqr_carver_new(guchar *buffer, gint width, gint height, gint channels)
{
    LqrCarver *r;
    if (!buffer || !engine_load()) return NULL;
    r = (LqrCarver *) calloc(1, sizeof(LqrCarver));
    if (!r) return NULL;
    r->eng = g_eng.carver_new(buffer, width, height, channels);
    if (!r->eng) {
        fprintf(stderr, "liblqr-1 (b200): %s\n", g_eng.last_error());
        free(r);
        return NULL;
    }
    r->progress = lqr_progress_new();
    if (!r->progress) {
        g_eng.carver_destroy(r->eng);
        free(r);
        return NULL;
    }
    r->resize_order = LQR_RES_ORDER_HOR;
    r->enl_step = 2.0f;
    r->channels = channels;
    r->session_update_step = 1;
    /* ownership of `buffer` passes to the carver (render.c:220-223 never frees it): the pixels now live
     * in HBM, so the host copy is released right away.  g_try_new == malloc on glib >= 2.46. */
    free(buffer);
    return r;
}

static void destroy_node(LqrCarver *r, gboolean destroy_engine)
{
    LqrCarverList *it, *itn;
    LqrVMapList *vl, *vln;
    for (it = r->attached; it; it = itn) {
        itn = it->next;
        destroy_node(it->current, FALSE); /* the engine frees attached carvers with their root */
        free(it);
    }
    for (vl = r->flushed_vs; vl; vl = vln) {
        vln = vl->next;
        lqr_vmap_destroy(vl->current);
        free(vl);
    }
    if (destroy_engine) g_eng.carver_destroy(r->eng);
    free(r->progress);
    free(r);
}

void lqr_carver_destroy(LqrCarver *r)
{
    if (!r) return;
    destroy_node(r, r->root == NULL);
}

LqrRetVal lqr_carver_init(LqrCarver *r, gint delta_x, gfloat rigidity)
{
    LQR_CATCH_F(r != NULL);
    return (LqrRetVal) g_eng.carver_init(r->eng, delta_x, rigidity);
}

LqrRetVal lqr_carver_attach(LqrCarver *r, LqrCarver *aux)
{
    LqrCarverList *node, **tail;
    LQR_CATCH_F(r != NULL && aux != NULL);
    LQR_CATCH((LqrRetVal) g_eng.carver_attach(r->eng, aux->eng));
    LQR_CATCH_MEM(node = (LqrCarverList *) malloc(sizeof(LqrCarverList)));
    node->current = aux;
    node->next = NULL;
    for (tail = &r->attached; *tail; tail = &(*tail)->next) {}
    *tail = node;
    aux->root = r;
    return LQR_OK;
}

/* ------------------------------------------------------------------ knobs */
LqrRetVal lqr_carver_set_energy_function_builtin(LqrCarver *r, LqrEnergyFuncBuiltinType ef)
{
    LQR_CATCH_F(r != NULL);
    return (LqrRetVal) g_eng.set_energy_function(r->eng, (int) ef);
}

void lqr_carver_set_resize_order(LqrCarver *r, LqrResizeOrder o)
{
    if (r) r->resize_order = o;
}

void lqr_carver_set_progress(LqrCarver *r, LqrProgress *p)
{
    if (!r) return;
    free(r->progress);
    r->progress = p;
}

void lqr_carver_set_side_switch_frequency(LqrCarver *r, guint f)
{
    if (r) g_eng.set_side_switch_frequency(r->eng, f);
}

LqrRetVal lqr_carver_set_enl_step(LqrCarver *r, gfloat s)
{
    LQR_CATCH_F(r != NULL);
    LQR_CATCH_F((s > 1) && (s <= 2));
    r->enl_step = s;
    return LQR_OK;
}

void lqr_carver_set_dump_vmaps(LqrCarver *r)
{
    if (r) r->dump_vmaps = TRUE;
}

void lqr_carver_set_no_dump_vmaps(LqrCarver *r)
{
    if (r) r->dump_vmaps = FALSE;
}

/* ------------------------------------------------------------------ getters */
gint lqr_carver_get_width(LqrCarver *r) { return EG(r, B200C_TRANSPOSED) ? EG(r, B200C_H) : EG(r, B200C_W); }
gint lqr_carver_get_height(LqrCarver *r) { return EG(r, B200C_TRANSPOSED) ? EG(r, B200C_W) : EG(r, B200C_H); }
gint lqr_carver_get_ref_width(LqrCarver *r)
{
    return EG(r, B200C_TRANSPOSED) ? EG(r, B200C_H_START) : EG(r, B200C_W_START);
}
gint lqr_carver_get_ref_height(LqrCarver *r)
{
    return EG(r, B200C_TRANSPOSED) ? EG(r, B200C_W_START) : EG(r, B200C_H_START);
}
gint lqr_carver_get_channels(LqrCarver *r) { return r->channels; }
gint lqr_carver_get_orientation(LqrCarver *r) { return EG(r, B200C_TRANSPOSED) ? 1 : 0; }
gint lqr_carver_get_depth(LqrCarver *r) { return EG(r, B200C_W0) - EG(r, B200C_W_START); }
gfloat lqr_carver_get_enl_step(LqrCarver *r) { return r->enl_step; }

/* ------------------------------------------------------------------ masks */
LqrRetVal lqr_carver_bias_add_rgb_area(LqrCarver *r, guchar *rgb, gint bias_factor, gint channels, gint width,
                                       gint height, gint x_off, gint y_off)
{
    LQR_CATCH_F(r != NULL && rgb != NULL); /* the plug-in does not NULL-check its mask buffer (io_functions.c:92) */
    invalidate_lines(r);
    return (LqrRetVal) g_eng.bias_add_rgb_area(r->eng, rgb, bias_factor, channels, width, height, x_off, y_off);
}

LqrRetVal lqr_carver_rigmask_add_rgb_area(LqrCarver *r, guchar *rgb, gint channels, gint width, gint height,
                                          gint x_off, gint y_off)
{
    LQR_CATCH_F(r != NULL && rgb != NULL);
    invalidate_lines(r);
    return (LqrRetVal) g_eng.rigmask_add_rgb_area(r->eng, rgb, channels, width, height, x_off, y_off);
}

/* ------------------------------------------------------------------ seam maps */
static LqrVMap *vmap_snapshot(LqrCarver *r)
{
    LqrVMap *v = (LqrVMap *) malloc(sizeof(LqrVMap));
    if (!v) return NULL;
    v->width = lqr_carver_get_ref_width(r);
    v->height = lqr_carver_get_ref_height(r);
    v->depth = lqr_carver_get_depth(r);
    v->orientation = lqr_carver_get_orientation(r);
    v->buffer = (gint *) malloc(sizeof(gint) * (size_t) v->width * v->height);
    if (!v->buffer || g_eng.vmap(r->eng, v->buffer) != B200C_OK) {
        free(v->buffer);
        free(v);
        return NULL;
    }
    return v;
}

LqrVMap *lqr_vmap_dump(LqrCarver *r) { return r ? vmap_snapshot(r) : NULL; }

static LqrRetVal vmap_internal_dump(LqrCarver *r)
{
    LqrVMapList *node, **tail;
    LqrVMap *v = vmap_snapshot(r);
    LQR_CATCH_MEM(v);
    node = (LqrVMapList *) malloc(sizeof(LqrVMapList));
    if (!node) {
        lqr_vmap_destroy(v);
        return LQR_NOMEM;
    }
    node->current = v;
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

/* ------------------------------------------------------------------ resize driver (A.10) */
static gint step_limit(gfloat enl_step, gint ref)
{
    gint d = (gint) ((enl_step - 1) * ref) - 1;
    return d < 1 ? 1 : d;
}

/* engine -> shim progress hook: called on the caller's thread at every progress point, once the device has completed
 * the seams it reports.  gimp_progress_update is cast to the hook's type (render.c:773): its gboolean TRUE is LQR_OK;
 * anything else cancels the resize, as in liblqr (LQR_USRCANCEL). */
static int on_seam(void *user, int seam_index)
{
    LqrCarver *r = (LqrCarver *) user;
    gint done = seam_index + r->session_rescale_current;
    if (done % r->session_update_step == 0 && r->progress && r->progress->update) {
        LqrRetVal ret = r->progress->update((gdouble) done / (gdouble) r->session_rescale_total);
        if (ret != LQR_OK) return 1;
    }
    return 0;
}

static LqrRetVal resize_direction(LqrCarver *r, gint target, gboolean along_w)
{
    gboolean need_flip = along_w ? EG(r, B200C_TRANSPOSED) : !EG(r, B200C_TRANSPOSED);
    gint ref = need_flip ? EG(r, B200C_H_START) : EG(r, B200C_W_START);
    gint cur = need_flip ? EG(r, B200C_H) : EG(r, B200C_W);
    gint delta = target - ref, gamma = target - cur, delta_max = step_limit(r->enl_step, ref);
    LqrProgress *p = r->progress;

    if (delta < 0) {
        delta = -delta;
        delta_max = delta;
    }
    r->session_rescale_total = gamma > 0 ? gamma : -gamma;
    r->session_rescale_current = 0;
    {
        gfloat step = r->session_rescale_total * (p ? p->update_step : 0.02f);
        r->session_update_step = (gint) (step > 1 ? step : 1);
    }
    if (r->session_rescale_total && p && p->init)
        p->init(along_w ? p->init_width_message : p->init_height_message);

    while (gamma) {
        gint delta0 = delta < delta_max ? delta : delta_max, new_w, w_start;
        delta -= delta0;
        if (along_w ? EG(r, B200C_TRANSPOSED) : !EG(r, B200C_TRANSPOSED))
            LQR_CATCH((LqrRetVal) g_eng.transpose(r->eng));
        w_start = EG(r, B200C_W_START);
        new_w = target < w_start + delta_max ? target : w_start + delta_max;
        gamma = target - new_w;
        LQR_CATCH((LqrRetVal) g_eng.build_maps_phase(r->eng, delta0 + 1, r->session_update_step,
                                                     r->session_rescale_current % r->session_update_step, on_seam, r));
        LQR_CATCH((LqrRetVal) g_eng.set_width(r->eng, new_w));
        r->session_rescale_current = r->session_rescale_total - (gamma > 0 ? gamma : -gamma);
        if (r->dump_vmaps) LQR_CATCH(vmap_internal_dump(r));
        if (new_w < target) {
            LQR_CATCH((LqrRetVal) g_eng.flatten(r->eng));
            delta_max = step_limit(r->enl_step, EG(r, B200C_W_START));
        }
    }
    if (r->session_rescale_total && p && p->end)
        p->end(along_w ? p->end_width_message : p->end_height_message);
    return LQR_OK;
}

LqrRetVal lqr_carver_resize(LqrCarver *r, gint w1, gint h1)
{
    LqrRetVal ret;
    LQR_CATCH_F(r != NULL);
    LQR_CATCH_F((w1 >= 1) && (h1 >= 1));
    LQR_CATCH_F(r->root == NULL);
    invalidate_lines(r);
    if (r->resize_order == LQR_RES_ORDER_HOR) {
        ret = resize_direction(r, w1, TRUE);
        if (ret == LQR_OK) ret = resize_direction(r, h1, FALSE);
    } else {
        ret = resize_direction(r, h1, FALSE);
        if (ret == LQR_OK) ret = resize_direction(r, w1, TRUE);
    }
    if (ret != LQR_OK) fprintf(stderr, "liblqr-1 (b200): resize failed: %s\n", g_eng.last_error());
    return ret;
}

/* ------------------------------------------------------------------ batch of independent images
 * The plug-in's batch use (batch/batch-gimp-lqr.scm:19-66: one image per call) as ONE call: the resize driver above
 * for n carvers in lockstep.  Carvers that agree in geometry and knobs share every launch of the engine (one host
 * thread, the images' row-serial chains side by side on the SMs); results are those of n separate lqr_carver_resize
 * calls.  Carvers that do not agree are resized one after the other. */
static gboolean batch_compatible(LqrCarver **rs, gint n)
{
    gint i;
    for (i = 0; i < n; i++) {
        LqrCarver *a = rs[0], *b = rs[i];
        if (!b || b->root) return FALSE;
        if (EG(a, B200C_W) != EG(b, B200C_W) || EG(a, B200C_H) != EG(b, B200C_H) ||
            EG(a, B200C_W_START) != EG(b, B200C_W_START) || EG(a, B200C_H_START) != EG(b, B200C_H_START) ||
            EG(a, B200C_W0) != EG(b, B200C_W0) || EG(a, B200C_H0) != EG(b, B200C_H0) ||
            EG(a, B200C_TRANSPOSED) != EG(b, B200C_TRANSPOSED) || EG(a, B200C_LEVEL) != EG(b, B200C_LEVEL) ||
            EG(a, B200C_MAX_LEVEL) != EG(b, B200C_MAX_LEVEL) || a->enl_step != b->enl_step ||
            a->resize_order != b->resize_order)
            return FALSE;
    }
    return TRUE;
}

static LqrRetVal resize_direction_batch(LqrCarver **rs, gint n, gint target, gboolean along_w, B200Carver **engs)
{
    LqrCarver *r = rs[0];
    gboolean need_flip = along_w ? EG(r, B200C_TRANSPOSED) : !EG(r, B200C_TRANSPOSED);
    gint ref = need_flip ? EG(r, B200C_H_START) : EG(r, B200C_W_START);
    gint cur = need_flip ? EG(r, B200C_H) : EG(r, B200C_W);
    gint delta = target - ref, gamma = target - cur, delta_max = step_limit(r->enl_step, ref), i;
    const gint total = gamma > 0 ? gamma : -gamma;

    if (delta < 0) {
        delta = -delta;
        delta_max = delta;
    }
    for (i = 0; i < n && total; i++) {
        LqrProgress *p = rs[i]->progress;
        if (p && p->init) p->init(along_w ? p->init_width_message : p->init_height_message);
    }
    while (gamma) {
        gint delta0 = delta < delta_max ? delta : delta_max, new_w, w_start;
        delta -= delta0;
        if (along_w ? EG(r, B200C_TRANSPOSED) : !EG(r, B200C_TRANSPOSED))
            for (i = 0; i < n; i++) LQR_CATCH((LqrRetVal) g_eng.transpose(rs[i]->eng));
        w_start = EG(r, B200C_W_START);
        new_w = target < w_start + delta_max ? target : w_start + delta_max;
        gamma = target - new_w;
        LQR_CATCH((LqrRetVal) g_eng.batch_build_maps(engs, n, delta0 + 1));
        for (i = 0; i < n; i++) {
            LQR_CATCH((LqrRetVal) g_eng.set_width(rs[i]->eng, new_w));
            if (rs[i]->dump_vmaps) LQR_CATCH(vmap_internal_dump(rs[i]));
        }
        if (new_w < target) {
            for (i = 0; i < n; i++) LQR_CATCH((LqrRetVal) g_eng.flatten(rs[i]->eng));
            delta_max = step_limit(r->enl_step, EG(r, B200C_W_START));
        }
    }
    for (i = 0; i < n && total; i++) {
        LqrProgress *p = rs[i]->progress;
        if (p && p->end) p->end(along_w ? p->end_width_message : p->end_height_message);
    }
    return LQR_OK;
}

LqrRetVal lqr_b200_batch_resize(LqrCarver **rs, gint n, gint w1, gint h1)
{
    LqrRetVal ret = LQR_OK;
    B200Carver **engs;
    gint i;
    LQR_CATCH_F(rs != NULL && n >= 1);
    LQR_CATCH_F((w1 >= 1) && (h1 >= 1));
    for (i = 0; i < n; i++) LQR_CATCH_F(rs[i] != NULL);
    if (n == 1 || !batch_compatible(rs, n)) {
        for (i = 0; i < n && ret == LQR_OK; i++) ret = lqr_carver_resize(rs[i], w1, h1);
        return ret;
    }
    engs = (B200Carver **) malloc(sizeof(B200Carver *) * (size_t) n);
    LQR_CATCH_MEM(engs);
    for (i = 0; i < n; i++) {
        engs[i] = rs[i]->eng;
        invalidate_lines(rs[i]);
    }
    if (rs[0]->resize_order == LQR_RES_ORDER_HOR) {
        ret = resize_direction_batch(rs, n, w1, TRUE, engs);
        if (ret == LQR_OK) ret = resize_direction_batch(rs, n, h1, FALSE, engs);
    } else {
        ret = resize_direction_batch(rs, n, h1, FALSE, engs);
        if (ret == LQR_OK) ret = resize_direction_batch(rs, n, w1, TRUE, engs);
    }
    free(engs);
    if (ret != LQR_OK) fprintf(stderr, "liblqr-1 (b200): batch resize failed: %s\n", g_eng.last_error());
    return ret;
}

LqrRetVal lqr_carver_flatten(LqrCarver *r)
{
    LQR_CATCH_F(r != NULL);
    invalidate_lines(r);
    return (LqrRetVal) g_eng.flatten(r->eng);
}

/* ------------------------------------------------------------------ read-out (A.12) */
static gboolean fetch_lines(LqrCarver *r)
{
    if (r->lines) return TRUE;
    if (g_eng.readout(r->eng, &r->lines) != B200C_OK) {
        fprintf(stderr, "liblqr-1 (b200): read-out failed: %s\n", g_eng.last_error());
        r->lines = NULL;
        return FALSE;
    }
    r->line_w = EG(r, B200C_W);
    r->line_h = EG(r, B200C_H);
    r->cur_line = 0;
    r->cur_x = 0;
    r->lines_ready = 0;
    return TRUE;
}

/* the row the cursor is on has arrived in the host buffer (waits for its chunk if it has not) */
static gboolean line_ready(LqrCarver *r, gint line)
{
    if (line < r->lines_ready) return TRUE;
    r->lines_ready = g_eng.readout_rows(r->eng, line);
    if (r->lines_ready <= line) {
        fprintf(stderr, "liblqr-1 (b200): read-out failed: %s\n", g_eng.last_error());
        r->lines = NULL;
        return FALSE;
    }
    return TRUE;
}

void lqr_carver_scan_reset(LqrCarver *r)
{
    if (!r) return;
    r->cur_line = 0;
    r->cur_x = 0;
}

gboolean lqr_carver_scan_by_row(LqrCarver *r) { return EG(r, B200C_TRANSPOSED) ? FALSE : TRUE; }

gboolean lqr_carver_scan_line(LqrCarver *r, gint *n, guchar **rgb)
{
    if (!r || !fetch_lines(r)) return FALSE;
    if (r->cur_line >= r->line_h) {
        lqr_carver_scan_reset(r);
        return FALSE;
    }
    if (!line_ready(r, r->cur_line)) return FALSE;
    *n = r->cur_line;
    *rgb = (guchar *) (r->lines + (size_t) r->cur_line * r->line_w * r->channels);
    r->cur_line++;
    r->cur_x = 0;
    return TRUE;
}

gboolean lqr_carver_scan(LqrCarver *r, gint *x, gint *y, guchar **rgb)
{
    gint k, transposed;
    const guchar *src;
    if (!r || !fetch_lines(r)) return FALSE;
    if (r->cur_line >= r->line_h) {
        lqr_carver_scan_reset(r);
        return FALSE;
    }
    if (!line_ready(r, r->cur_line)) return FALSE;
    transposed = EG(r, B200C_TRANSPOSED);
    *x = transposed ? r->cur_line : r->cur_x;
    *y = transposed ? r->cur_x : r->cur_line;
    src = r->lines + ((size_t) r->cur_line * r->line_w + r->cur_x) * r->channels;
    for (k = 0; k < r->channels; k++) r->pixel[k] = src[k];
    *rgb = r->pixel;
    if (++r->cur_x >= r->line_w) {
        r->cur_x = 0;
        r->cur_line++;
    }
    return TRUE;
}

LqrRetVal lqr_carver_get_true_energy(LqrCarver *r, gfloat *buffer, gint orientation)
{
    LQR_CATCH_F(r != NULL && buffer != NULL);
    LQR_CATCH_F(orientation == 0 || orientation == 1);
    LQR_CATCH_F(r->root == NULL);
    invalidate_lines(r);
    if (EG(r, B200C_W) != EG(r, B200C_W_START) - EG(r, B200C_MAX_LEVEL) + 1)
        LQR_CATCH((LqrRetVal) g_eng.flatten(r->eng));
    if (orientation != lqr_carver_get_orientation(r)) LQR_CATCH((LqrRetVal) g_eng.transpose(r->eng));
    return (LqrRetVal) g_eng.true_energy(r->eng, buffer);
}

/* ------------------------------------------------------------------ attached-carver list */
LqrCarverList *lqr_carver_list_start(LqrCarver *r) { return r->attached; }
LqrCarver *lqr_carver_list_current(LqrCarverList *l) { return l ? l->current : NULL; }
LqrCarverList *lqr_carver_list_next(LqrCarverList *l) { return l ? l->next : NULL; }

/* ------------------------------------------------------------------ test hook: engine handle of a carver */
LQR_PUBLIC void *lqr_b200_engine_handle(LqrCarver *r) { return r ? (void *) r->eng : NULL; }
