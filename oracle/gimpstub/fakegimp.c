/* fakegimp.c -- an in-memory GIMP for the reference's render path (test infrastructure, oracle/_ref).
 *
 * The reference plug-in's own object code (src/render.c, src/io_functions.c, compiled unmodified) talks to libgimp;
 * here images and layers are plain buffers in this process.  Semantics follow libgimp 2.8 where the render path depends
 * on them: gimp_layer_resize(w, h, offx, offy) places the old content at (offx, offy) of the new extent and moves the
 * layer's offsets by (-offx, -offy); shadow pixel regions are merged by gimp_drawable_merge_shadow; a GRAY image asked
 * for seam maps is converted to RGB first (render.c:161-168).  The fg_* entry points are the test's side door.
 */
#include <libgimp/gimp.h>
#include <stdint.h>

#define FG_MAX 256
#define FG_PUBLIC __attribute__((visibility("default")))

typedef struct {
    int used, image, w, h, bpp, xoff, yoff, visible, lock_alpha, inserted;
    guchar *px, *shadow;
    char name[256];
} FgLayer;
typedef struct {
    int used, w, h, base_type, active;
} FgImage;

static FgLayer g_layer[FG_MAX];
static FgImage g_image[FG_MAX];
static long g_progress_init, g_progress_update, g_progress_end, g_messages;
static double g_last_fraction;
static char g_last_message[512];

static FgLayer *L(gint32 id) { return (id > 0 && id < FG_MAX && g_layer[id].used) ? &g_layer[id] : NULL; }
static FgImage *I(gint32 id) { return (id > 0 && id < FG_MAX && g_image[id].used) ? &g_image[id] : NULL; }

void g_message(const gchar *format, ...)
{
    va_list ap;
    va_start(ap, format);
    vsnprintf(g_last_message, sizeof g_last_message, format, ap);
    va_end(ap);
    g_messages++;
}

static gint32 layer_alloc(int image, int w, int h, int bpp, const char *name)
{
    for (int id = 1; id < FG_MAX; ++id)
        if (!g_layer[id].used) {
            FgLayer *l = &g_layer[id];
            memset(l, 0, sizeof *l);
            l->used = 1, l->image = image, l->w = w, l->h = h, l->bpp = bpp, l->visible = 1;
            l->px = (guchar *) calloc((size_t) w * h * bpp + 1, 1);
            snprintf(l->name, sizeof l->name, "%s", name ? name : "layer");
            return l->px ? id : -1;
        }
    return -1;
}

/* ---- the test's side door --------------------------------------------------------------------------------- */
FG_PUBLIC void fg_reset(void)
{
    for (int i = 0; i < FG_MAX; ++i) {
        free(g_layer[i].px);
        free(g_layer[i].shadow);
    }
    memset(g_layer, 0, sizeof g_layer);
    memset(g_image, 0, sizeof g_image);
    g_progress_init = g_progress_update = g_progress_end = g_messages = 0;
    g_last_message[0] = 0;
}
FG_PUBLIC int fg_image_new(int w, int h, int base_type) { return gimp_image_new(w, h, (GimpImageBaseType) base_type); }
FG_PUBLIC int fg_layer_add(int image, int w, int h, int bpp, int xoff, int yoff, const unsigned char *pixels, const char *name)
{
    const gint32 id = layer_alloc(image, w, h, bpp, name);
    if (id < 0) return -1;
    memcpy(g_layer[id].px, pixels, (size_t) w * h * bpp);
    g_layer[id].xoff = xoff, g_layer[id].yoff = yoff, g_layer[id].inserted = 1;
    if (I(image) && !I(image)->active) I(image)->active = id;
    return id;
}
FG_PUBLIC int fg_layer_info(int id, int out[6])
{
    FgLayer *l = L(id);
    if (!l) return 0;
    out[0] = l->w, out[1] = l->h, out[2] = l->bpp, out[3] = l->xoff, out[4] = l->yoff, out[5] = l->image;
    return 1;
}
FG_PUBLIC const unsigned char *fg_layer_pixels(int id) { return L(id) ? L(id)->px : NULL; }
FG_PUBLIC const char *fg_layer_name(int id) { return L(id) ? L(id)->name : NULL; }
/* ids of the layers of `image` in creation order */
FG_PUBLIC int fg_image_layers(int image, int *ids, int cap)
{
    int n = 0;
    for (int id = 1; id < FG_MAX; ++id)
        if (g_layer[id].used && g_layer[id].image == image && g_layer[id].inserted && n < cap) ids[n++] = id;
    return n;
}
FG_PUBLIC void fg_progress_counts(long out[4])
{
    out[0] = g_progress_init, out[1] = g_progress_update, out[2] = g_progress_end, out[3] = g_messages;
}
FG_PUBLIC const char *fg_last_message(void) { return g_last_message; }

/* ---- images ------------------------------------------------------------------------------------------------ */
gboolean gimp_image_is_valid(gint32 id) { return I(id) != NULL; }
gboolean gimp_drawable_is_valid(gint32 id) { return L(id) != NULL; }
gint32 gimp_image_get_active_layer(gint32 id) { return I(id) ? I(id)->active : -1; }
gboolean gimp_image_set_active_layer(gint32 id, gint32 layer)
{
    if (!I(id)) return FALSE;
    I(id)->active = layer;
    return TRUE;
}
gboolean gimp_image_unset_active_channel(gint32 id) { (void) id; return TRUE; }
GimpImageBaseType gimp_image_base_type(gint32 id) { return I(id) ? (GimpImageBaseType) I(id)->base_type : GIMP_RGB; }
gboolean gimp_image_convert_rgb(gint32 id)
{
    if (!I(id)) return FALSE;
    for (int k = 1; k < FG_MAX; ++k) {
        FgLayer *l = &g_layer[k];
        if (!l->used || l->image != id || l->bpp > 2) continue;
        const int nb = l->bpp + 2;
        guchar *np = (guchar *) calloc((size_t) l->w * l->h * nb + 1, 1);
        for (size_t i = 0; i < (size_t) l->w * l->h; ++i) {
            np[i * nb] = np[i * nb + 1] = np[i * nb + 2] = l->px[i * l->bpp];
            if (l->bpp == 2) np[i * nb + 3] = l->px[i * 2 + 1];
        }
        free(l->px);
        l->px = np, l->bpp = nb;
    }
    I(id)->base_type = GIMP_RGB;
    return TRUE;
}
gint32 gimp_image_new(gint w, gint h, GimpImageBaseType type)
{
    for (int id = 1; id < FG_MAX; ++id)
        if (!g_image[id].used) {
            g_image[id].used = 1, g_image[id].w = w, g_image[id].h = h, g_image[id].base_type = type, g_image[id].active = 0;
            return id;
        }
    return -1;
}
gboolean gimp_image_insert_layer(gint32 image, gint32 layer, gint32 parent, gint position)
{
    (void) parent, (void) position;
    if (!I(image) || !L(layer)) return FALSE;
    L(layer)->image = image, L(layer)->inserted = 1;
    return TRUE;
}
gboolean gimp_image_resize(gint32 id, gint w, gint h, gint offx, gint offy)
{
    if (!I(id)) return FALSE;
    I(id)->w = w, I(id)->h = h;
    for (int k = 1; k < FG_MAX; ++k)
        if (g_layer[k].used && g_layer[k].image == id) g_layer[k].xoff += offx, g_layer[k].yoff += offy;
    return TRUE;
}
gboolean gimp_image_undo_group_start(gint32 id) { (void) id; return TRUE; }
gboolean gimp_image_undo_group_end(gint32 id) { (void) id; return TRUE; }
gint32 gimp_display_new(gint32 id) { (void) id; return 1; }

/* ---- selection / masks: the fake image has neither ------------------------------------------------------- */
gboolean gimp_layer_is_floating_sel(gint32 id) { (void) id; return FALSE; }
gboolean gimp_floating_sel_to_layer(gint32 id) { (void) id; return TRUE; }
gint32 gimp_layer_get_mask(gint32 id) { (void) id; return -1; }
gboolean gimp_layer_remove_mask(gint32 id, GimpMaskApplyMode mode) { (void) id, (void) mode; return TRUE; }
gboolean gimp_selection_is_empty(gint32 id) { (void) id; return TRUE; }
gint32 gimp_selection_save(gint32 id) { (void) id; return -1; }
gboolean gimp_selection_none(gint32 id) { (void) id; return TRUE; }

/* ---- layers ------------------------------------------------------------------------------------------------ */
gint32 gimp_layer_new(gint32 image, const gchar *name, gint w, gint h, GimpImageType type, gdouble opacity, GimpLayerModeEffects mode)
{
    (void) opacity, (void) mode;
    static const int bpp_of[4] = {3, 4, 1, 2};
    const gint32 id = layer_alloc(image, w, h, bpp_of[type & 3], name);
    return id;
}
static gint32 layer_dup(gint32 src, gint32 image)
{
    FgLayer *s = L(src);
    if (!s) return -1;
    const gint32 id = layer_alloc(image, s->w, s->h, s->bpp, s->name);
    if (id < 0) return -1;
    memcpy(g_layer[id].px, s->px, (size_t) s->w * s->h * s->bpp);
    g_layer[id].xoff = s->xoff, g_layer[id].yoff = s->yoff, g_layer[id].lock_alpha = s->lock_alpha;
    return id;
}
gint32 gimp_layer_copy(gint32 id) { return L(id) ? layer_dup(id, L(id)->image) : -1; }
gint32 gimp_layer_new_from_drawable(gint32 id, gint32 dest_image) { return layer_dup(id, dest_image); }
gboolean gimp_layer_resize(gint32 id, gint nw, gint nh, gint offx, gint offy)
{
    FgLayer *l = L(id);
    if (!l || nw < 1 || nh < 1) return FALSE;
    guchar *np = (guchar *) calloc((size_t) nw * nh * l->bpp + 1, 1); /* new area is transparent */
    if (!np) return FALSE;
    for (int y = 0; y < nh; ++y) {
        const int sy = y - offy;
        if (sy < 0 || sy >= l->h) continue;
        int x0 = offx > 0 ? offx : 0, x1 = l->w + offx < nw ? l->w + offx : nw;
        if (x1 > x0) memcpy(np + ((size_t) y * nw + x0) * l->bpp, l->px + ((size_t) sy * l->w + (x0 - offx)) * l->bpp, (size_t) (x1 - x0) * l->bpp);
    }
    free(l->px);
    free(l->shadow);
    l->shadow = NULL;
    l->px = np, l->w = nw, l->h = nh, l->xoff -= offx, l->yoff -= offy;
    return TRUE;
}
gboolean gimp_layer_resize_to_image_size(gint32 id)
{
    FgLayer *l = L(id);
    if (!l || !I(l->image)) return FALSE;
    return gimp_layer_resize(id, I(l->image)->w, I(l->image)->h, l->xoff, l->yoff);
}
gboolean gimp_layer_scale(gint32 id, gint nw, gint nh, gboolean local_origin)
{
    (void) local_origin;
    FgLayer *l = L(id); /* nearest neighbour: the standard scale-back modes are outside the engine's path */
    if (!l || nw < 1 || nh < 1) return FALSE;
    guchar *np = (guchar *) calloc((size_t) nw * nh * l->bpp + 1, 1);
    if (!np) return FALSE;
    for (int y = 0; y < nh; ++y)
        for (int x = 0; x < nw; ++x)
            memcpy(np + ((size_t) y * nw + x) * l->bpp, l->px + ((size_t) (y * l->h / nh) * l->w + (x * l->w / nw)) * l->bpp, l->bpp);
    free(l->px);
    free(l->shadow);
    l->shadow = NULL;
    l->px = np, l->w = nw, l->h = nh;
    return TRUE;
}
gboolean gimp_layer_translate(gint32 id, gint dx, gint dy)
{
    if (!L(id)) return FALSE;
    L(id)->xoff += dx, L(id)->yoff += dy;
    return TRUE;
}
gboolean gimp_layer_get_lock_alpha(gint32 id) { return L(id) ? L(id)->lock_alpha : FALSE; }
gboolean gimp_layer_set_lock_alpha(gint32 id, gboolean v)
{
    if (!L(id)) return FALSE;
    L(id)->lock_alpha = v;
    return TRUE;
}

/* ---- drawables --------------------------------------------------------------------------------------------- */
gint gimp_drawable_width(gint32 id) { return L(id) ? L(id)->w : 0; }
gint gimp_drawable_height(gint32 id) { return L(id) ? L(id)->h : 0; }
gint gimp_drawable_bpp(gint32 id) { return L(id) ? L(id)->bpp : 0; }
gboolean gimp_drawable_offsets(gint32 id, gint *x, gint *y)
{
    if (!L(id)) return FALSE;
    *x = L(id)->xoff, *y = L(id)->yoff;
    return TRUE;
}
gchar *gimp_drawable_get_name(gint32 id) { return L(id) ? L(id)->name : (gchar *) ""; }
gboolean gimp_drawable_set_name(gint32 id, const gchar *name)
{
    if (!L(id)) return FALSE;
    snprintf(L(id)->name, sizeof L(id)->name, "%s", name);
    return TRUE;
}
gboolean gimp_drawable_set_visible(gint32 id, gboolean v)
{
    if (!L(id)) return FALSE;
    L(id)->visible = v;
    return TRUE;
}
gboolean gimp_drawable_fill(gint32 id, GimpFillType fill)
{
    FgLayer *l = L(id);
    if (!l) return FALSE;
    memset(l->px, fill == GIMP_WHITE_FILL ? 255 : 0, (size_t) l->w * l->h * l->bpp);
    return TRUE;
}
GimpDrawable *gimp_drawable_get(gint32 id)
{
    FgLayer *l = L(id);
    if (!l) return NULL;
    GimpDrawable *d = (GimpDrawable *) calloc(1, sizeof *d);
    d->drawable_id = id, d->width = l->w, d->height = l->h, d->bpp = l->bpp;
    return d;
}
void gimp_drawable_detach(GimpDrawable *d) { free(d); }
void gimp_drawable_flush(GimpDrawable *d) { (void) d; }
gboolean gimp_drawable_merge_shadow(gint32 id, gboolean undo)
{
    (void) undo;
    FgLayer *l = L(id);
    if (!l || !l->shadow) return FALSE;
    memcpy(l->px, l->shadow, (size_t) l->w * l->h * l->bpp);
    free(l->shadow);
    l->shadow = NULL;
    return TRUE;
}
gboolean gimp_drawable_update(gint32 id, gint x, gint y, gint w, gint h) { (void) id, (void) x, (void) y, (void) w, (void) h; return TRUE; }

/* ---- pixel regions ------------------------------------------------------------------------------------------ */
void gimp_pixel_rgn_init(GimpPixelRgn *pr, GimpDrawable *d, gint x, gint y, gint w, gint h, gint dirty, gint shadow)
{
    memset(pr, 0, sizeof *pr);
    pr->drawable = d, pr->bpp = d->bpp, pr->x = x, pr->y = y, pr->w = w, pr->h = h, pr->dirty = dirty, pr->shadow = shadow;
    FgLayer *l = L(d->drawable_id);
    if (l && shadow && !l->shadow) {
        l->shadow = (guchar *) malloc((size_t) l->w * l->h * l->bpp + 1);
        memcpy(l->shadow, l->px, (size_t) l->w * l->h * l->bpp);
    }
}
static guchar *rgn_base(GimpPixelRgn *pr, int for_write)
{
    FgLayer *l = L(pr->drawable->drawable_id);
    if (!l) return NULL;
    return (for_write && pr->shadow) ? l->shadow : l->px;
}
void gimp_pixel_rgn_get_row(GimpPixelRgn *pr, guchar *buf, gint x, gint y, gint width)
{
    FgLayer *l = L(pr->drawable->drawable_id);
    const guchar *b = rgn_base(pr, 0);
    if (!l || !b || y < 0 || y >= l->h || x < 0 || x + width > l->w) return;
    memcpy(buf, b + ((size_t) y * l->w + x) * l->bpp, (size_t) width * l->bpp);
}
void gimp_pixel_rgn_set_row(GimpPixelRgn *pr, const guchar *buf, gint x, gint y, gint width)
{
    FgLayer *l = L(pr->drawable->drawable_id);
    guchar *b = rgn_base(pr, 1);
    if (!l || !b || y < 0 || y >= l->h || x < 0 || x + width > l->w) return;
    memcpy(b + ((size_t) y * l->w + x) * l->bpp, buf, (size_t) width * l->bpp);
}
void gimp_pixel_rgn_set_col(GimpPixelRgn *pr, const guchar *buf, gint x, gint y, gint height)
{
    FgLayer *l = L(pr->drawable->drawable_id);
    guchar *b = rgn_base(pr, 1);
    if (!l || !b || x < 0 || x >= l->w || y < 0 || y + height > l->h) return;
    for (int j = 0; j < height; ++j) memcpy(b + ((size_t) (y + j) * l->w + x) * l->bpp, buf + (size_t) j * l->bpp, l->bpp);
}

/* ---- progress, tiles, colours ---------------------------------------------------------------------------------- */
gboolean gimp_progress_init(const gchar *message) { (void) message; g_progress_init++; return TRUE; }
gboolean gimp_progress_update(gdouble f) { g_last_fraction = f; g_progress_update++; return TRUE; }
gboolean gimp_progress_end(void) { g_progress_end++; return TRUE; }
guint gimp_tile_width(void) { return 64; }
guint gimp_tile_height(void) { return 64; }
void gimp_tile_cache_size(unsigned long kb) { (void) kb; }
void gimp_rgba_set(GimpRGB *c, gdouble r, gdouble g, gdouble b, gdouble a) { c->r = r, c->g = g, c->b = b, c->a = a; }
