/* lqr.h -- drop-in public header for the LqrCarver C API, B200-native build.
 *
 * This header is the boundary of the hot path: it declares every entry point
 * that gimp-lqr-plugin's render path binds in the external liblqr library
 * (`#include <lqr.h>`, reference src/render.c:25, src/io_functions.c:23) plus
 * the few extra liblqr calls the north-star names (lqr_carver_scan,
 * energy read-back).  The plug-in treats every Lqr* type as opaque, so only
 * names, argument meaning, enum values and ownership rules are contractual.
 *
 * Each declaration cites the reference call site it serves (file:line under
 * the reference tree).  Two libraries export exactly this API:
 *   - liblqr-1.so          the plain-C shim that dlopen()s libb200carve.so
 *                          (the CUDA engine) -- the product;
 *   - liblqr_oracle.so     the single-threaded CPU restatement in oracle/
 *                          (test infrastructure only).
 *
 * glib is optional: when <glib.h> is not available the g* typedefs below are
 * used (identical ABI: gint=int, guchar=unsigned char, gfloat=float,
 * gdouble=double, gboolean=int, gpointer=void*).
 */
#ifndef __LQR_H__
#define __LQR_H__ /* guard name required by reference src/io_functions.h:22-24 */

#if defined(LQR_USE_GLIB)
#include <glib.h>
#elif defined(__has_include)
#if __has_include(<glib.h>)
#include <glib.h>
#define LQR_USE_GLIB 1
#endif
#endif

#ifndef LQR_USE_GLIB
#ifndef __G_TYPES_H__
typedef int gint;
typedef unsigned int guint;
typedef unsigned char guchar;
typedef char gchar;
typedef float gfloat;
typedef double gdouble;
typedef int gboolean;
typedef void *gpointer;
typedef int gint32;
#endif
#ifndef TRUE
#define TRUE 1
#endif
#ifndef FALSE
#define FALSE 0
#endif
#endif /* !LQR_USE_GLIB */

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define LQR_PUBLIC __attribute__((visibility("default")))
#else
#define LQR_PUBLIC
#endif

#define LQR_MAX_NAME_LENGTH (1024) /* sizes stack buffers at render.c:114-115, io_functions.c:296 */
#define LQR_PROGRESS_MAX_MESSAGE_LENGTH (1024)

/* ---- return values (render.c:42-46 tests == LQR_NOMEM; LQR_OK must be 1 because
 *      gboolean-returning gimp_progress_* are cast to the callback types, render.c:772-773) */
typedef enum _LqrRetVal {
    LQR_ERROR = 0,
    LQR_OK = 1,
    LQR_NOMEM = 2,
    LQR_USRCANCEL = 3
} LqrRetVal;

/* ---- legacy error macros used by the plug-in (io_functions.c:46,94,125,247) */
#define LQR_CATCH(expr) do { LqrRetVal lqr_ret_val__; \
    if ((lqr_ret_val__ = (expr)) != LQR_OK) { return lqr_ret_val__; } } while (0)
#define LQR_CATCH_F(expr) do { if ((expr) == FALSE) { return LQR_ERROR; } } while (0)
#define LQR_CATCH_MEM(expr) do { if ((expr) == NULL) { return LQR_NOMEM; } } while (0)
#define LQR_TRY_N_N(assign) do { if ((assign) == NULL) { return NULL; } } while (0)
#ifndef LQR_DISABLE_LEGACY_MACROS
#define CATCH(expr) LQR_CATCH(expr)
#define CATCH_F(expr) LQR_CATCH_F(expr)
#define CATCH_MEM(expr) LQR_CATCH_MEM(expr)
#define TRY_N_N(assign) LQR_TRY_N_N(assign)
#endif

/* ---- enums whose numeric values cross the PDB boundary (main.c:77-78,545; batch-gimp-lqr.scm:51-52) */
typedef enum _LqrResizeOrder {
    LQR_RES_ORDER_HOR = 0,
    LQR_RES_ORDER_VERT = 1
} LqrResizeOrder;

typedef enum _LqrEnergyFuncBuiltinType {
    LQR_EF_GRAD_NORM = 0,
    LQR_EF_GRAD_SUMABS = 1,
    LQR_EF_GRAD_XABS = 2,
    LQR_EF_LUMA_GRAD_NORM = 3,
    LQR_EF_LUMA_GRAD_SUMABS = 4,
    LQR_EF_LUMA_GRAD_XABS = 5,
    LQR_EF_NULL = 6
} LqrEnergyFuncBuiltinType; /* the seven combo-box entries of interface.c:2138-2145 */

typedef enum _LqrImageType {
    LQR_RGB_IMAGE = 0,
    LQR_RGBA_IMAGE = 1,
    LQR_GREY_IMAGE = 2,
    LQR_GREYA_IMAGE = 3
} LqrImageType;

/* ---- opaque handles */
typedef struct _LqrCarver LqrCarver;
typedef struct _LqrCarverList LqrCarverList;
typedef struct _LqrVMap LqrVMap;
typedef struct _LqrVMapList LqrVMapList;
typedef struct _LqrProgress LqrProgress;

typedef LqrRetVal (*LqrProgressFuncInit)(const gchar *init_message);
typedef LqrRetVal (*LqrProgressFuncUpdate)(gdouble percentage);
typedef LqrRetVal (*LqrProgressFuncEnd)(const gchar *end_message);
typedef LqrRetVal (*LqrVMapFunc)(LqrVMap *vmap, gpointer data);

/* ---- carver life cycle */
/* render.c:222,894 -- ADOPTS `buffer` (freed in lqr_carver_destroy); channels 1..4, alpha last for 2/4 */
LQR_PUBLIC LqrCarver *lqr_carver_new(guchar *buffer, gint width, gint height, gint channels);
/* render.c:376, interface_I.c:427 -- also destroys attached carvers, flushed vmaps, progress */
LQR_PUBLIC void lqr_carver_destroy(LqrCarver *r);
/* render.c:224 -- allocates the seam-search state; rigidity_map[dx] = rigidity*|dx|^1.5/h */
LQR_PUBLIC LqrRetVal lqr_carver_init(LqrCarver *r, gint delta_x, gfloat rigidity);
/* render.c:897 -- aux carver (same size) shares the root's visibility map; ADOPTS aux */
LQR_PUBLIC LqrRetVal lqr_carver_attach(LqrCarver *r, LqrCarver *aux);

/* ---- the hot path */
/* render.c:318,328,529 */
LQR_PUBLIC LqrRetVal lqr_carver_resize(LqrCarver *r, gint w1, gint h1);
/* render.c:325,636 */
LQR_PUBLIC LqrRetVal lqr_carver_flatten(LqrCarver *r);

/* ---- masks (buffers are BORROWED: plug-in frees them, io_functions.c:97,128) */
/* io_functions.c:94-95 */
LQR_PUBLIC LqrRetVal lqr_carver_bias_add_rgb_area(LqrCarver *r, guchar *rgb, gint bias_factor, gint channels,
                                                  gint width, gint height, gint x_off, gint y_off);
/* io_functions.c:125-126 */
LQR_PUBLIC LqrRetVal lqr_carver_rigmask_add_rgb_area(LqrCarver *r, guchar *rgb, gint channels,
                                                     gint width, gint height, gint x_off, gint y_off);

/* ---- knobs (render.c:234-241) */
LQR_PUBLIC LqrRetVal lqr_carver_set_energy_function_builtin(LqrCarver *r, LqrEnergyFuncBuiltinType ef_ind);
LQR_PUBLIC void lqr_carver_set_resize_order(LqrCarver *r, LqrResizeOrder resize_order);
LQR_PUBLIC void lqr_carver_set_progress(LqrCarver *r, LqrProgress *p); /* ADOPTS p */
LQR_PUBLIC void lqr_carver_set_side_switch_frequency(LqrCarver *r, guint switch_frequency);
LQR_PUBLIC LqrRetVal lqr_carver_set_enl_step(LqrCarver *r, gfloat enl_step); /* valid in (1,2] */
LQR_PUBLIC void lqr_carver_set_dump_vmaps(LqrCarver *r);
LQR_PUBLIC void lqr_carver_set_no_dump_vmaps(LqrCarver *r);

/* ---- getters (io_functions.c:145,168; render.c:49,547-551,654-658) */
LQR_PUBLIC gint lqr_carver_get_width(LqrCarver *r);
LQR_PUBLIC gint lqr_carver_get_height(LqrCarver *r);
LQR_PUBLIC gint lqr_carver_get_ref_width(LqrCarver *r);
LQR_PUBLIC gint lqr_carver_get_ref_height(LqrCarver *r);
LQR_PUBLIC gint lqr_carver_get_channels(LqrCarver *r);
LQR_PUBLIC gint lqr_carver_get_orientation(LqrCarver *r); /* 0 = rows are image rows, 1 = transposed */
LQR_PUBLIC gint lqr_carver_get_depth(LqrCarver *r);       /* seams available = w0 - w_start */
LQR_PUBLIC gfloat lqr_carver_get_enl_step(LqrCarver *r);

/* ---- read-out (io_functions.c:155-164): engine-owned line buffer, valid until the next scan call */
LQR_PUBLIC gboolean lqr_carver_scan_line(LqrCarver *r, gint *n, guchar **rgb);
LQR_PUBLIC gboolean lqr_carver_scan_by_row(LqrCarver *r);
LQR_PUBLIC gboolean lqr_carver_scan(LqrCarver *r, gint *x, gint *y, guchar **rgb); /* pixel-wise (north-star) */
LQR_PUBLIC void lqr_carver_scan_reset(LqrCarver *r);

/* ---- energy read-back: the hook the 1e-6 energy parity check uses.
 *      buffer is width*height floats in image orientation; orientation 0 = horizontal seams search
 *      (vertical seams), 1 = transposed. */
LQR_PUBLIC LqrRetVal lqr_carver_get_true_energy(LqrCarver *r, gfloat *buffer, gint orientation);

/* ---- attached-carver list (render.c:370,839-841,912-914) */
LQR_PUBLIC LqrCarverList *lqr_carver_list_start(LqrCarver *r);
LQR_PUBLIC LqrCarver *lqr_carver_list_current(LqrCarverList *list);
LQR_PUBLIC LqrCarverList *lqr_carver_list_next(LqrCarverList *list);

/* ---- visibility (seam) maps (render.c:344,725,747; io_functions.c:216-219,312) */
LQR_PUBLIC LqrVMap *lqr_vmap_dump(LqrCarver *r); /* caller-owned */
LQR_PUBLIC void lqr_vmap_destroy(LqrVMap *vmap);
LQR_PUBLIC gint *lqr_vmap_get_data(LqrVMap *vmap);
LQR_PUBLIC gint lqr_vmap_get_width(LqrVMap *vmap);
LQR_PUBLIC gint lqr_vmap_get_height(LqrVMap *vmap);
LQR_PUBLIC gint lqr_vmap_get_depth(LqrVMap *vmap);
LQR_PUBLIC gint lqr_vmap_get_orientation(LqrVMap *vmap);
LQR_PUBLIC LqrVMapList *lqr_vmap_list_start(LqrCarver *r); /* engine-owned */
LQR_PUBLIC LqrVMap *lqr_vmap_list_current(LqrVMapList *list);
LQR_PUBLIC LqrVMapList *lqr_vmap_list_next(LqrVMapList *list);
LQR_PUBLIC LqrRetVal lqr_vmap_list_foreach(LqrVMapList *list, LqrVMapFunc func, gpointer data);

/* ---- progress reporting (render.c:767-779): callbacks run synchronously on the caller's thread */
LQR_PUBLIC LqrProgress *lqr_progress_new(void);
LQR_PUBLIC LqrRetVal lqr_progress_set_init(LqrProgress *p, LqrProgressFuncInit init_func);
LQR_PUBLIC LqrRetVal lqr_progress_set_update(LqrProgress *p, LqrProgressFuncUpdate update_func);
LQR_PUBLIC LqrRetVal lqr_progress_set_end(LqrProgress *p, LqrProgressFuncEnd end_func);
LQR_PUBLIC LqrRetVal lqr_progress_set_update_step(LqrProgress *p, gfloat update_step);
LQR_PUBLIC LqrRetVal lqr_progress_set_init_width_message(LqrProgress *p, const gchar *message);
LQR_PUBLIC LqrRetVal lqr_progress_set_init_height_message(LqrProgress *p, const gchar *message);
LQR_PUBLIC LqrRetVal lqr_progress_set_end_width_message(LqrProgress *p, const gchar *message);
LQR_PUBLIC LqrRetVal lqr_progress_set_end_height_message(LqrProgress *p, const gchar *message);

#ifdef __cplusplus
}
#endif

#endif /* __LQR_H__ */
