/* libgimp/gimp.h -- an in-memory stand-in for the part of libgimp the reference's render path calls
 * (src/render.c, src/io_functions.c): images and layers live in this process (fakegimp.c), pixel regions read and write
 * them row by row exactly as the plug-in asks.  Test infrastructure (oracle/_ref), not product code. */
#ifndef __LIBGIMP_STUB_H__
#define __LIBGIMP_STUB_H__
#include <glib.h>

typedef enum { GIMP_RGB = 0, GIMP_GRAY = 1, GIMP_INDEXED = 2 } GimpImageBaseType;
typedef enum { GIMP_RGB_IMAGE = 0, GIMP_RGBA_IMAGE = 1, GIMP_GRAY_IMAGE = 2, GIMP_GRAYA_IMAGE = 3 } GimpImageType;
typedef enum { GIMP_NORMAL_MODE = 0 } GimpLayerModeEffects;
typedef enum { GIMP_FOREGROUND_FILL = 0, GIMP_BACKGROUND_FILL, GIMP_WHITE_FILL, GIMP_TRANSPARENT_FILL } GimpFillType;
typedef enum { GIMP_MASK_APPLY = 0, GIMP_MASK_DISCARD = 1 } GimpMaskApplyMode;

typedef struct { gdouble r, g, b, a; } GimpRGB;
typedef struct { gint32 drawable_id; guint width, height, bpp; } GimpDrawable;
typedef struct {
    guchar *data;
    GimpDrawable *drawable;
    gint bpp, rowstride, x, y, w, h;
    guint dirty, shadow;
    gint process_count;
} GimpPixelRgn;

gboolean gimp_image_is_valid(gint32 image_ID);
gboolean gimp_drawable_is_valid(gint32 drawable_ID);
gint32 gimp_image_get_active_layer(gint32 image_ID);
gboolean gimp_image_set_active_layer(gint32 image_ID, gint32 layer_ID);
gboolean gimp_image_unset_active_channel(gint32 image_ID);
GimpImageBaseType gimp_image_base_type(gint32 image_ID);
gboolean gimp_image_convert_rgb(gint32 image_ID);
gint32 gimp_image_new(gint width, gint height, GimpImageBaseType type);
gboolean gimp_image_insert_layer(gint32 image_ID, gint32 layer_ID, gint32 parent_ID, gint position);
gboolean gimp_image_resize(gint32 image_ID, gint new_width, gint new_height, gint offx, gint offy);
gboolean gimp_image_undo_group_start(gint32 image_ID);
gboolean gimp_image_undo_group_end(gint32 image_ID);
gint32 gimp_display_new(gint32 image_ID);

gboolean gimp_layer_is_floating_sel(gint32 layer_ID);
gboolean gimp_floating_sel_to_layer(gint32 layer_ID);
gint32 gimp_layer_get_mask(gint32 layer_ID);
gboolean gimp_layer_remove_mask(gint32 layer_ID, GimpMaskApplyMode mode);
gboolean gimp_selection_is_empty(gint32 image_ID);
gint32 gimp_selection_save(gint32 image_ID);
gboolean gimp_selection_none(gint32 image_ID);

gint32 gimp_layer_new(gint32 image_ID, const gchar *name, gint width, gint height, GimpImageType type, gdouble opacity,
                      GimpLayerModeEffects mode);
gint32 gimp_layer_copy(gint32 layer_ID);
gint32 gimp_layer_new_from_drawable(gint32 drawable_ID, gint32 dest_image_ID);
gboolean gimp_layer_resize(gint32 layer_ID, gint new_width, gint new_height, gint offx, gint offy);
gboolean gimp_layer_resize_to_image_size(gint32 layer_ID);
gboolean gimp_layer_scale(gint32 layer_ID, gint new_width, gint new_height, gboolean local_origin);
gboolean gimp_layer_translate(gint32 layer_ID, gint offx, gint offy);
gboolean gimp_layer_get_lock_alpha(gint32 layer_ID);
gboolean gimp_layer_set_lock_alpha(gint32 layer_ID, gboolean lock_alpha);

gint gimp_drawable_width(gint32 drawable_ID);
gint gimp_drawable_height(gint32 drawable_ID);
gint gimp_drawable_bpp(gint32 drawable_ID);
gboolean gimp_drawable_offsets(gint32 drawable_ID, gint *offset_x, gint *offset_y);
gchar *gimp_drawable_get_name(gint32 drawable_ID);
gboolean gimp_drawable_set_name(gint32 drawable_ID, const gchar *name);
gboolean gimp_drawable_set_visible(gint32 drawable_ID, gboolean visible);
gboolean gimp_drawable_fill(gint32 drawable_ID, GimpFillType fill_type);
GimpDrawable *gimp_drawable_get(gint32 drawable_ID);
void gimp_drawable_detach(GimpDrawable *drawable);
void gimp_drawable_flush(GimpDrawable *drawable);
gboolean gimp_drawable_merge_shadow(gint32 drawable_ID, gboolean undo);
gboolean gimp_drawable_update(gint32 drawable_ID, gint x, gint y, gint width, gint height);

void gimp_pixel_rgn_init(GimpPixelRgn *pr, GimpDrawable *drawable, gint x, gint y, gint width, gint height, gint dirty,
                         gint shadow);
void gimp_pixel_rgn_get_row(GimpPixelRgn *pr, guchar *buf, gint x, gint y, gint width);
void gimp_pixel_rgn_set_row(GimpPixelRgn *pr, const guchar *buf, gint x, gint y, gint width);
void gimp_pixel_rgn_set_col(GimpPixelRgn *pr, const guchar *buf, gint x, gint y, gint height);

gboolean gimp_progress_init(const gchar *message);
gboolean gimp_progress_update(gdouble percentage);
gboolean gimp_progress_end(void);
guint gimp_tile_width(void);
guint gimp_tile_height(void);
void gimp_tile_cache_size(unsigned long kilobytes);
void gimp_rgba_set(GimpRGB *rgba, gdouble r, gdouble g, gdouble b, gdouble a);
#endif
