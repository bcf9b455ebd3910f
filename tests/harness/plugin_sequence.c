/* plugin_sequence.c -- libgimp-free replay of the plug-in's render path against any library exporting the
 * LqrCarver API (the product liblqr-1.so or the CPU oracle), resolved with dlopen/dlsym exactly like a plug-in
 * binary linked with $(LQR_LIBS) would resolve them.
 *
 * The call order, the buffer ownership and the per-line read-out loop are those of the reference:
 *   render_init_carver      src/render.c:220-248   (rgb_buffer_from_layer -> lqr_carver_new -> init -> masks -> knobs)
 *   render_noninteractive   src/render.c:318-376   (resize [-> flatten -> resize] -> vmaps -> scan loop -> destroy)
 *   write_carver_to_layer   src/io_functions.c:155-164 (scan_line; set_row when scan_by_row else set_col)
 * The gimp_pixel_rgn_* calls are replaced by memcpy into caller-provided host buffers.
 *
 * Build: gcc -O2 -fPIC -shared -I../../include plugin_sequence.c -ldl -o libplugin_sequence.so
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <malloc.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "lqr.h"

typedef struct {
    int width, height, bpp;       /* layer geometry (gimp_drawable_width/height/bpp) */
    int new_width, new_height;    /* PlugInVals.new_width / new_height */
    int pres_coeff, disc_coeff;   /* main.c:66,68 */
    float rigidity;               /* main.c:69 */
    int delta_x;                  /* main.c:71 */
    float enl_step;               /* main.c:72, percent */
    int nrg_func, res_order;      /* main.c:77-78 */
    int output_seams;             /* main.c:76 */
    int scaleback;                /* LQRBACK, render.c:320-329 */
    int no_disc_on_enlarge;       /* main.c:82 */
    int mask_bpp;                 /* bpp of the pres / disc / rigmask layers (same size as the layer), 0 = none */
    int resize_aux_layers;        /* attach the mask layers as aux carvers (render.c:243-248) */
} HarnessVals;

typedef struct {
    int out_width, out_height;
    int n_vmaps, vmap_width, vmap_height, vmap_depth; /* first flushed seam map */
    int n_progress_updates;
    double ms_new, ms_setup, ms_resize, ms_scan, ms_total;
} HarnessResult;

typedef struct {
    void *dl;
    LqrCarver *(*carver_new)(guchar *, gint, gint, gint);
    void (*carver_destroy)(LqrCarver *);
    LqrRetVal (*carver_init)(LqrCarver *, gint, gfloat);
    LqrRetVal (*carver_attach)(LqrCarver *, LqrCarver *);
    LqrRetVal (*carver_resize)(LqrCarver *, gint, gint);
    LqrRetVal (*carver_flatten)(LqrCarver *);
    LqrRetVal (*bias_add_rgb_area)(LqrCarver *, guchar *, gint, gint, gint, gint, gint, gint);
    LqrRetVal (*rigmask_add_rgb_area)(LqrCarver *, guchar *, gint, gint, gint, gint, gint);
    LqrRetVal (*set_energy_function_builtin)(LqrCarver *, LqrEnergyFuncBuiltinType);
    void (*set_resize_order)(LqrCarver *, LqrResizeOrder);
    void (*set_progress)(LqrCarver *, LqrProgress *);
    void (*set_side_switch_frequency)(LqrCarver *, guint);
    LqrRetVal (*set_enl_step)(LqrCarver *, gfloat);
    void (*set_dump_vmaps)(LqrCarver *);
    gint (*get_width)(LqrCarver *);
    gint (*get_height)(LqrCarver *);
    gboolean (*scan_line)(LqrCarver *, gint *, guchar **);
    gboolean (*scan_by_row)(LqrCarver *);
    LqrVMapList *(*vmap_list_start)(LqrCarver *);
    LqrRetVal (*vmap_list_foreach)(LqrVMapList *, LqrVMapFunc, gpointer);
    gint *(*vmap_get_data)(LqrVMap *);
    gint (*vmap_get_width)(LqrVMap *);
    gint (*vmap_get_height)(LqrVMap *);
    gint (*vmap_get_depth)(LqrVMap *);
    LqrProgress *(*progress_new)(void);
    LqrRetVal (*progress_set_init)(LqrProgress *, LqrProgressFuncInit);
    LqrRetVal (*progress_set_update)(LqrProgress *, LqrProgressFuncUpdate);
    LqrRetVal (*progress_set_end)(LqrProgress *, LqrProgressFuncEnd);
    LqrRetVal (*progress_set_init_width_message)(LqrProgress *, const gchar *);
    LqrRetVal (*progress_set_init_height_message)(LqrProgress *, const gchar *);
} Api;

#define SYM(field, name)                                                   \
    do {                                                                   \
        *(void **) (&api->field) = dlsym(api->dl, name);                   \
        if (!api->field) {                                                 \
            fprintf(stderr, "plugin_sequence: %s lacks %s\n", path, name); \
            return 0;                                                      \
        }                                                                  \
    } while (0)

static int bind_api(Api *api, const char *path)
{
    api->dl = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!api->dl) {
        fprintf(stderr, "plugin_sequence: dlopen(%s): %s\n", path, dlerror());
        return 0;
    }
    SYM(carver_new, "lqr_carver_new");
    SYM(carver_destroy, "lqr_carver_destroy");
    SYM(carver_init, "lqr_carver_init");
    SYM(carver_attach, "lqr_carver_attach");
    SYM(carver_resize, "lqr_carver_resize");
    SYM(carver_flatten, "lqr_carver_flatten");
    SYM(bias_add_rgb_area, "lqr_carver_bias_add_rgb_area");
    SYM(rigmask_add_rgb_area, "lqr_carver_rigmask_add_rgb_area");
    SYM(set_energy_function_builtin, "lqr_carver_set_energy_function_builtin");
    SYM(set_resize_order, "lqr_carver_set_resize_order");
    SYM(set_progress, "lqr_carver_set_progress");
    SYM(set_side_switch_frequency, "lqr_carver_set_side_switch_frequency");
    SYM(set_enl_step, "lqr_carver_set_enl_step");
    SYM(set_dump_vmaps, "lqr_carver_set_dump_vmaps");
    SYM(get_width, "lqr_carver_get_width");
    SYM(get_height, "lqr_carver_get_height");
    SYM(scan_line, "lqr_carver_scan_line");
    SYM(scan_by_row, "lqr_carver_scan_by_row");
    SYM(vmap_list_start, "lqr_vmap_list_start");
    SYM(vmap_list_foreach, "lqr_vmap_list_foreach");
    SYM(vmap_get_data, "lqr_vmap_get_data");
    SYM(vmap_get_width, "lqr_vmap_get_width");
    SYM(vmap_get_height, "lqr_vmap_get_height");
    SYM(vmap_get_depth, "lqr_vmap_get_depth");
    SYM(progress_new, "lqr_progress_new");
    SYM(progress_set_init, "lqr_progress_set_init");
    SYM(progress_set_update, "lqr_progress_set_update");
    SYM(progress_set_end, "lqr_progress_set_end");
    SYM(progress_set_init_width_message, "lqr_progress_set_init_width_message");
    SYM(progress_set_init_height_message, "lqr_progress_set_init_height_message");
    return 1;
}

/* The plug-in lives in a long-running host whose allocator recycles the layer-sized pixel buffers (render.c:220 fills a
 * fresh g_try_new buffer per run).  glibc would serve such a block (31.6 MiB at 4K RGBA) from a fresh mapping and
 * unmap it on free -- ~8000 page faults per run, bimodal by allocation history; keep freed blocks in the heap
 * instead so that runs are repeatable.  Applies to this test / bench process only. */
__attribute__((constructor)) static void harness_allocator_setup(void)
{
    mallopt(M_MMAP_THRESHOLD, 32 << 20);
    mallopt(M_TRIM_THRESHOLD, 1 << 30);
}

static double now_ms(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

/* stand-ins for gimp_progress_init / gimp_progress_update / gimp_progress_end: gboolean TRUE == LQR_OK */
static int g_updates;
static LqrRetVal fake_progress_init(const gchar *m) { (void) m; return (LqrRetVal) TRUE; }
static LqrRetVal fake_progress_update(gdouble f) { (void) f; g_updates++; return (LqrRetVal) TRUE; }
static LqrRetVal fake_progress_end(const gchar *m) { (void) m; return (LqrRetVal) TRUE; }

/* rgb_buffer_from_layer (io_functions.c:29-68): a fresh g_try_new buffer filled row by row */
static guchar *rgb_buffer_from_layer(const unsigned char *layer, int w, int h, int bpp)
{
    guchar *buffer = (guchar *) malloc((size_t) bpp * w * h);
    int y;
    if (!buffer) return NULL;
    for (y = 0; y < h; y++) memcpy(buffer + (size_t) y * w * bpp, layer + (size_t) y * w * bpp, (size_t) w * bpp);
    return buffer;
}

/* write_carver_to_layer (io_functions.c:134-182) into a dense out buffer of out_w x out_h x bpp */
static int write_carver_to_layer(const Api *api, LqrCarver *r, unsigned char *out, int w, int h, int bpp)
{
    gint y;
    guchar *line;
    int n = 0;
    while (api->scan_line(r, &y, &line)) {
        if (api->scan_by_row(r)) {
            memcpy(out + (size_t) y * w * bpp, line, (size_t) w * bpp); /* gimp_pixel_rgn_set_row */
        } else {
            int i; /* gimp_pixel_rgn_set_col */
            for (i = 0; i < h; i++) memcpy(out + ((size_t) i * w + y) * bpp, line + (size_t) i * bpp, bpp);
        }
        n++;
    }
    return n;
}

typedef struct {
    const Api *api;
    int *dst;
    HarnessResult *res;
} VmapSink;

static LqrRetVal vmap_sink(LqrVMap *vmap, gpointer data) /* write_vmap_to_layer, io_functions.c:184-290 */
{
    VmapSink *s = (VmapSink *) data;
    if (s->res->n_vmaps == 0) {
        s->res->vmap_width = s->api->vmap_get_width(vmap);
        s->res->vmap_height = s->api->vmap_get_height(vmap);
        s->res->vmap_depth = s->api->vmap_get_depth(vmap);
        if (s->dst)
            memcpy(s->dst, s->api->vmap_get_data(vmap), sizeof(int) * (size_t) s->res->vmap_width * s->res->vmap_height);
    }
    s->res->n_vmaps++;
    return LQR_OK;
}

/* Returns 1 on success.  layer: width*height*bpp; pres / disc / rigmask: same size, mask_bpp each, or NULL.
 * out: at least max(width,new_width)*max(height,new_height)*bpp bytes; vmap_out: width*height ints or NULL. */
int harness_render(const char *liblqr_path, const unsigned char *layer, const HarnessVals *v, const unsigned char *pres,
                   const unsigned char *disc, const unsigned char *rigmask, unsigned char *out, int *vmap_out,
                   HarnessResult *res)
{
    Api api_storage, *api = &api_storage;
    LqrCarver *carver;
    LqrProgress *progress;
    guchar *rgb_buffer;
    const unsigned char *aux_src[3];
    int ignore_disc = 0, k;
    float rigidity;
    double t0, t1, t2, t3, t4;
    int new_w = v->new_width, new_h = v->new_height;

    memset(res, 0, sizeof *res);
    if (!bind_api(api, liblqr_path)) return 0;
    g_updates = 0;
    t0 = now_ms();

    rigidity = rigmask ? 3 * v->rigidity : v->rigidity;                        /* render.c:781-792 */
    if (v->no_disc_on_enlarge) {                                                /* render.c:794-821 */
        if (v->res_order == LQR_RES_ORDER_HOR)
            ignore_disc = new_w > v->width || (new_w == v->width && new_h > v->height);
        else
            ignore_disc = new_h > v->height || (new_h == v->height && new_w > v->width);
    }
    progress = api->progress_new();                                            /* render.c:211, 767-779 */
    if (!progress) return 0;
    api->progress_set_init(progress, fake_progress_init);
    api->progress_set_update(progress, fake_progress_update);
    api->progress_set_end(progress, fake_progress_end);
    api->progress_set_init_width_message(progress, "Resizing width...");
    api->progress_set_init_height_message(progress, "Resizing height...");

    rgb_buffer = rgb_buffer_from_layer(layer, v->width, v->height, v->bpp);   /* render.c:220 */
    if (!rgb_buffer) return 0;
    carver = api->carver_new(rgb_buffer, v->width, v->height, v->bpp);        /* render.c:222, adopts rgb_buffer */
    if (!carver) return 0;
    t1 = now_ms();
    if (api->carver_init(carver, v->delta_x, rigidity) != LQR_OK) return 0;   /* render.c:224 */
    if (pres && v->pres_coeff != 0) {                                          /* update_bias, io_functions.c:70-100 */
        guchar *rgb = rgb_buffer_from_layer(pres, v->width, v->height, v->mask_bpp);
        if (api->bias_add_rgb_area(carver, rgb, v->pres_coeff, v->mask_bpp, v->width, v->height, 0, 0) != LQR_OK) return 0;
        free(rgb);
    }
    if (disc && !ignore_disc && v->disc_coeff != 0) {
        guchar *rgb = rgb_buffer_from_layer(disc, v->width, v->height, v->mask_bpp);
        if (api->bias_add_rgb_area(carver, rgb, -v->disc_coeff, v->mask_bpp, v->width, v->height, 0, 0) != LQR_OK) return 0;
        free(rgb);
    }
    if (rigmask) {                                                             /* set_rigmask, io_functions.c:102-131 */
        guchar *rgb = rgb_buffer_from_layer(rigmask, v->width, v->height, v->mask_bpp);
        if (api->rigmask_add_rgb_area(carver, rgb, v->mask_bpp, v->width, v->height, 0, 0) != LQR_OK) return 0;
        free(rgb);
    }
    api->set_energy_function_builtin(carver, (LqrEnergyFuncBuiltinType) v->nrg_func); /* render.c:234 */
    api->set_resize_order(carver, (LqrResizeOrder) v->res_order);             /* render.c:235 */
    api->set_progress(carver, progress);                                      /* render.c:236, adopts progress */
    api->set_side_switch_frequency(carver, 2);                                /* render.c:237 */
    api->set_enl_step(carver, v->enl_step / 100);                             /* render.c:238 */
    if (v->output_seams) api->set_dump_vmaps(carver);                         /* render.c:239-242 */
    aux_src[0] = pres, aux_src[1] = disc, aux_src[2] = rigmask;
    if (v->resize_aux_layers)                                                 /* attach_aux_carver, render.c:881-900 */
        for (k = 0; k < 3; k++)
            if (aux_src[k]) {
                guchar *rgb = rgb_buffer_from_layer(aux_src[k], v->width, v->height, v->mask_bpp);
                LqrCarver *aux = api->carver_new(rgb, v->width, v->height, v->mask_bpp);
                if (!aux || api->carver_attach(carver, aux) != LQR_OK) return 0;
            }
    t2 = now_ms();

    if (api->carver_resize(carver, new_w, new_h) == LQR_NOMEM) return 0;     /* render.c:318 (MEM_CHECK1) */
    if (v->scaleback) {                                                        /* LQRBACK, render.c:324-329 */
        if (api->carver_flatten(carver) == LQR_NOMEM) return 0;
        new_w = v->width;
        new_h = v->height;
        if (api->carver_resize(carver, new_w, new_h) == LQR_NOMEM) return 0;
    }
    t3 = now_ms();
    if (v->output_seams) {                                                     /* write_all_vmaps, render.c:340-346 */
        VmapSink sink;
        sink.api = api;
        sink.dst = vmap_out;
        sink.res = res;
        api->vmap_list_foreach(api->vmap_list_start(carver), vmap_sink, &sink);
    }
    res->out_width = api->get_width(carver);
    res->out_height = api->get_height(carver);
    write_carver_to_layer(api, carver, out, res->out_width, res->out_height, v->bpp); /* render.c:366 */
    api->carver_destroy(carver);                                              /* render.c:376 */
    t4 = now_ms();

    res->n_progress_updates = g_updates;
    res->ms_new = t1 - t0;
    res->ms_setup = t2 - t1;
    res->ms_resize = t3 - t2;
    res->ms_scan = t4 - t3;
    res->ms_total = t4 - t0;
    return 1;
}

/* ---- a batch of independent layers, `nthreads` of them in flight: what a batch host (batch/batch-gimp-lqr.scm run over a
 * directory) does with one plug-in run per image.  One host thread per image in flight; every run is
 * the call sequence above.  outs: NULL (results dropped) or one output buffer per layer.  sums[0..4] receive the per-phase milliseconds summed over the images, *wall_ms the wall time. */
#include <pthread.h>

typedef struct {
    const char *path;
    const unsigned char *const *layers;
    unsigned char *const *outs; /* NULL, or one buffer per layer */
    const HarnessVals *v;
    int n, next, failed;
    double sums[5];
    pthread_mutex_t mu;
} BatchJob;

static void *batch_worker(void *arg)
{
    BatchJob *job = (BatchJob *) arg;
    const HarnessVals *v = job->v;
    const size_t ow = (size_t) (v->width > v->new_width ? v->width : v->new_width);
    const size_t oh = (size_t) (v->height > v->new_height ? v->height : v->new_height);
    unsigned char *out = (unsigned char *) malloc(ow * oh * v->bpp);
    double sums[5] = {0, 0, 0, 0, 0};
    int failed = out == NULL;
    while (!failed) {
        HarnessResult res;
        int i;
        pthread_mutex_lock(&job->mu);
        i = job->next < job->n ? job->next++ : -1;
        pthread_mutex_unlock(&job->mu);
        if (i < 0) break;
        if (!harness_render(job->path, job->layers[i], v, NULL, NULL, NULL, job->outs ? job->outs[i] : out, NULL, &res) ||
            res.out_width != v->new_width || res.out_height != v->new_height)
            failed = 1;
        sums[0] += res.ms_new, sums[1] += res.ms_setup, sums[2] += res.ms_resize, sums[3] += res.ms_scan, sums[4] += res.ms_total;
    }
    free(out);
    pthread_mutex_lock(&job->mu);
    job->failed |= failed;
    for (int k = 0; k < 5; k++) job->sums[k] += sums[k];
    pthread_mutex_unlock(&job->mu);
    return NULL;
}

int harness_render_batch(const char *liblqr_path, const unsigned char *const *layers, unsigned char *const *outs, int n,
                         int nthreads, const HarnessVals *v, double *sums, double *wall_ms)
{
    BatchJob job;
    pthread_t th[64];
    double t0;
    int k;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 64) nthreads = 64;
    memset(&job, 0, sizeof job);
    job.path = liblqr_path, job.layers = layers, job.outs = outs, job.v = v, job.n = n;
    pthread_mutex_init(&job.mu, NULL);
    t0 = now_ms();
    for (k = 0; k < nthreads; k++)
        if (pthread_create(&th[k], NULL, batch_worker, &job) != 0) {
            nthreads = k;
            job.failed = 1;
            break;
        }
    for (k = 0; k < nthreads; k++) pthread_join(th[k], NULL);
    *wall_ms = now_ms() - t0;
    for (k = 0; k < 5; k++) sums[k] = job.sums[k];
    pthread_mutex_destroy(&job.mu);
    return !job.failed;
}

/* ---- the same batch, LOCKSTEP: the images are taken in groups of `group`; a group's carvers are set up one after the
 * other (the call sequence above, no masks), resized by ONE call -- lqr_b200_batch_resize, which advances all of them
 * with shared launches on the device -- and written back one after the other.  `nthreads` groups are in flight, so the
 * uploads and read-outs of one group overlap the seams of another. */
typedef LqrRetVal (*BatchResizeFn)(LqrCarver **, gint, gint, gint);

typedef struct {
    BatchJob job;
    int group;
} LockstepJob;

static void *lockstep_worker(void *arg)
{
    LockstepJob *lj = (LockstepJob *) arg;
    BatchJob *job = &lj->job;
    const HarnessVals *v = job->v;
    const size_t ow = (size_t) (v->width > v->new_width ? v->width : v->new_width);
    const size_t oh = (size_t) (v->height > v->new_height ? v->height : v->new_height);
    Api api_storage, *api = &api_storage;
    BatchResizeFn batch_resize;
    unsigned char *scratch = (unsigned char *) malloc(ow * oh * v->bpp);
    LqrCarver **cs = (LqrCarver **) calloc((size_t) lj->group, sizeof(LqrCarver *));
    double sums[5] = {0, 0, 0, 0, 0};
    int failed = scratch == NULL || cs == NULL || !bind_api(api, job->path);
    batch_resize = failed ? NULL : (BatchResizeFn) dlsym(api->dl, "lqr_b200_batch_resize");
    while (!failed) {
        int i0, g, k;
        double t0, t1, t2, t3;
        pthread_mutex_lock(&job->mu);
        i0 = job->next;
        g = job->n - i0 < lj->group ? job->n - i0 : lj->group;
        job->next += g > 0 ? g : 0;
        pthread_mutex_unlock(&job->mu);
        if (g <= 0) break;
        t0 = now_ms();
        for (k = 0; k < g && !failed; k++) {
            guchar *rgb = rgb_buffer_from_layer(job->layers[i0 + k], v->width, v->height, v->bpp);      /* render.c:220 */
            cs[k] = rgb ? api->carver_new(rgb, v->width, v->height, v->bpp) : NULL;                      /* render.c:222 */
            if (!cs[k] || api->carver_init(cs[k], v->delta_x, v->rigidity) != LQR_OK) {                  /* render.c:224 */
                failed = 1;
                break;
            }
            api->set_energy_function_builtin(cs[k], (LqrEnergyFuncBuiltinType) v->nrg_func);             /* render.c:234 */
            api->set_resize_order(cs[k], (LqrResizeOrder) v->res_order);                                 /* render.c:235 */
            api->set_side_switch_frequency(cs[k], 2);                                                    /* render.c:237 */
            api->set_enl_step(cs[k], v->enl_step / 100);                                                 /* render.c:238 */
        }
        t1 = now_ms();
        if (!failed) {
            if (batch_resize) {
                if (batch_resize(cs, g, v->new_width, v->new_height) != LQR_OK) failed = 1;
            } else {
                for (k = 0; k < g; k++)
                    if (api->carver_resize(cs[k], v->new_width, v->new_height) != LQR_OK) failed = 1;   /* render.c:318 */
            }
        }
        t2 = now_ms();
        for (k = 0; k < g; k++) {
            if (!cs[k]) continue;
            if (!failed) {
                if (api->get_width(cs[k]) != v->new_width || api->get_height(cs[k]) != v->new_height) failed = 1;
                write_carver_to_layer(api, cs[k], job->outs ? job->outs[i0 + k] : scratch, v->new_width, v->new_height, v->bpp);
            }
            api->carver_destroy(cs[k]);                                                                   /* render.c:376 */
            cs[k] = NULL;
        }
        t3 = now_ms();
        sums[0] += t1 - t0, sums[2] += t2 - t1, sums[3] += t3 - t2, sums[4] += t3 - t0;
    }
    free(scratch);
    free(cs);
    pthread_mutex_lock(&job->mu);
    job->failed |= failed;
    for (int k = 0; k < 5; k++) job->sums[k] += sums[k];
    pthread_mutex_unlock(&job->mu);
    return NULL;
}

int harness_render_lockstep(const char *liblqr_path, const unsigned char *const *layers, unsigned char *const *outs, int n,
                            int group, int nthreads, const HarnessVals *v, double *sums, double *wall_ms)
{
    LockstepJob lj;
    pthread_t th[64];
    double t0;
    int k;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 64) nthreads = 64;
    if (group < 1) group = 1;
    memset(&lj, 0, sizeof lj);
    lj.job.path = liblqr_path, lj.job.layers = layers, lj.job.outs = outs, lj.job.v = v, lj.job.n = n;
    lj.group = group;
    pthread_mutex_init(&lj.job.mu, NULL);
    t0 = now_ms();
    for (k = 0; k < nthreads; k++)
        if (pthread_create(&th[k], NULL, lockstep_worker, &lj) != 0) {
            nthreads = k;
            lj.job.failed = 1;
            break;
        }
    for (k = 0; k < nthreads; k++) pthread_join(th[k], NULL);
    *wall_ms = now_ms() - t0;
    for (k = 0; k < 5; k++) sums[k] = lj.job.sums[k];
    pthread_mutex_destroy(&lj.job.mu);
    return !lj.job.failed;
}
