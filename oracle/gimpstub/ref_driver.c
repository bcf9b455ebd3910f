/* ref_driver.c -- calls the reference plug-in's own render path the way its run() does for a non-interactive
 * invocation (reference src/main.c:413-444): render_init_carver + render_noninteractive, both from the reference's
 * UNMODIFIED src/render.c (and src/io_functions.c underneath), compiled here against include/lqr.h and the in-memory
 * libgimp of fakegimp.c.  Test infrastructure (oracle/_ref): it proves the header and the library are a drop-in for
 * the object code of the plug-in, and it is the reference's own loop for write_vmap_to_layer / write_carver_to_layer.
 */
#include "config.h"

#include <gtk/gtk.h>
#include <libgimp/gimp.h>
#include <lqr.h>

#include "io_functions.h"
#include "plugin-intl.h"
#include "main.h"
#include "render.h"

#define REF_PUBLIC __attribute__((visibility("default")))

/* iv: new_width, new_height, pres_layer_ID, pres_coeff, disc_layer_ID, disc_coeff, rigmask_layer_ID, delta_x,
 *     resize_aux_layers, resize_canvas, output_target, output_seams, nrg_func, res_order, mask_behavior, scaleback,
 *     scaleback_mode, no_disc_on_enlarge (the PlugInVals of src/main_common.h:34-60, in order, minus the floats);
 * fv: rigidity, enl_step; col: r1 g1 b1 r2 g2 b2 (PlugInColVals).  Returns render_success; *out_layer / *out_image
 * receive the drawable and image the result was written to. */
REF_PUBLIC int ref_run_noninteractive(int image_ID, int layer_ID, const int *iv, const float *fv, const double *col,
                                      int *out_image, int *out_layer)
{
    PlugInVals vals;
    PlugInImageVals image_vals;
    PlugInDrawableVals drawable_vals;
    PlugInColVals col_vals;
    CarverData *carver_data;
    gboolean render_success = FALSE;

    memset(&vals, 0, sizeof vals);
    vals.new_width = iv[0], vals.new_height = iv[1];
    vals.pres_layer_ID = iv[2], vals.pres_coeff = iv[3];
    vals.disc_layer_ID = iv[4], vals.disc_coeff = iv[5];
    vals.rigidity = fv[0];
    vals.rigmask_layer_ID = iv[6];
    vals.delta_x = iv[7];
    vals.enl_step = fv[1];
    vals.resize_aux_layers = iv[8], vals.resize_canvas = iv[9];
    vals.output_target = iv[10], vals.output_seams = iv[11];
    vals.nrg_func = iv[12], vals.res_order = iv[13], vals.mask_behavior = iv[14];
    vals.scaleback = iv[15], vals.scaleback_mode = iv[16], vals.no_disc_on_enlarge = iv[17];
    col_vals.r1 = col[0], col_vals.g1 = col[1], col_vals.b1 = col[2];
    col_vals.r2 = col[3], col_vals.g2 = col[4], col_vals.b2 = col[5];
    image_vals.image_ID = image_ID;
    drawable_vals.layer_ID = layer_ID;

    gimp_image_undo_group_start(image_ID);
    carver_data = render_init_carver(&image_vals, &drawable_vals, &vals, FALSE);
    if (carver_data) {
        image_vals.image_ID = carver_data->image_ID;
        drawable_vals.layer_ID = carver_data->layer_ID;
        if (image_ID != image_vals.image_ID) {
            gimp_image_undo_group_end(image_ID);
            image_ID = image_vals.image_ID;
            gimp_image_undo_group_start(image_ID);
        }
        render_success = render_noninteractive(&vals, &col_vals, carver_data);
    }
    gimp_image_undo_group_end(image_ID);
    *out_image = image_vals.image_ID;
    *out_layer = drawable_vals.layer_ID;
    return render_success;
}

/* the reference's seam-map colouring on its own (src/io_functions.c:184-290), given a library vmap handle */
REF_PUBLIC int ref_write_vmap(void *vmap, int image_ID, const char *name, int x_off, int y_off, const double *col, int *layer_io)
{
    VMapFuncArg data;
    gint32 id = *layer_io;
    data.image_ID = image_ID;
    data.name = (gchar *) name;
    data.x_off = x_off, data.y_off = y_off;
    gimp_rgba_set(&data.colour_start, col[0], col[1], col[2], 1);
    gimp_rgba_set(&data.colour_end, col[3], col[4], col[5], 1);
    data.vmap_layer_ID_p = &id;
    const LqrRetVal r = write_vmap_to_layer((LqrVMap *) vmap, &data);
    *layer_io = id;
    return r;
}
