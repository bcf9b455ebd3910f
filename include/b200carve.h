/* b200carve.h -- C ABI of the CUDA seam-carving engine (libb200carve.so, sm_100a).
 *
 * This is the library the plain-C shim (liblqr-1.so, include/lqr.h) dlopen()s.  Every entry point
 * takes plain pointers and sizes; no torch or C++ types cross the boundary.  Each function names the
 * liblqr operation it stands in for and the plug-in call site that reaches it (reference file:line);
 * liblqr itself is external to the reference tree (configure.ac:67-70), its semantics are restated in
 * SURVEY.md Appendix A (cited as A.n).
 *
 * All device state of a carver lives in HBM for the life of the handle (interactive mode keeps one
 * carver alive across many resize calls, interface_I.c:401-427).  Return convention: 1 = ok,
 * 0 = error, 2 = out of memory (host or device), 3 = cancelled -- the LqrRetVal values, so the shim
 * forwards them unchanged.  There is no CPU fallback: without a CUDA device every call fails.
 */
#ifndef B200CARVE_H
#define B200CARVE_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define B200C_API __attribute__((visibility("default")))
#else
#define B200C_API
#endif

#define B200C_ABI_VERSION 2
#define B200C_ERROR 0
#define B200C_OK 1
#define B200C_NOMEM 2
#define B200C_CANCEL 3

typedef struct B200Carver B200Carver;

/* called for the progress point "seam_index seams of this build_maps call are done" (liblqr calls its update hook
 * before it searches seam `seam_index`); nonzero = cancel */
typedef int (*b200c_progress_fn)(void *user, int seam_index);

B200C_API int b200c_abi_version(void);
B200C_API const char *b200c_last_error(void);
B200C_API int b200c_device_count(void);
B200C_API int b200c_set_device(int device); /* device used by carvers created afterwards on this thread */
B200C_API int b200c_device_cc(int device);  /* compute capability major * 10 + minor (100 = sm_100), -1 on error */

/* lqr_carver_new (render.c:222,894): copies the 8-bit interleaved image to HBM.  The host buffer stays
 * the caller's (the shim frees it; ownership rule of the Lqr API is the shim's business). */
B200C_API B200Carver *b200c_carver_new(const unsigned char *rgb, int width, int height, int channels);
/* same, image already resident in HBM (bench.py `value` leg); copied device-to-device on `stream` (may be NULL) */
B200C_API B200Carver *b200c_carver_new_device(const void *d_rgb, int width, int height, int channels);
B200C_API void b200c_carver_destroy(B200Carver *c); /* lqr_carver_destroy (render.c:376); frees attached carvers too */

/* lqr_carver_init (render.c:224, A.1): raw index table, en/m/least maps, rigidity map */
B200C_API int b200c_carver_init(B200Carver *c, int delta_x, float rigidity);
/* lqr_carver_attach (render.c:897, A.14): aux shares the root's visibility map; adopts aux */
B200C_API int b200c_carver_attach(B200Carver *root, B200Carver *aux);
/* lqr_carver_set_energy_function_builtin (render.c:234, A.3), ef in 0..6 */
B200C_API int b200c_carver_set_energy_function(B200Carver *c, int ef);
/* lqr_carver_set_side_switch_frequency (render.c:237, A.7) */
B200C_API int b200c_carver_set_side_switch_frequency(B200Carver *c, unsigned int frequency);
/* lqr_carver_bias_add_rgb_area / lqr_carver_rigmask_add_rgb_area (io_functions.c:94,125; A.4); mask is a host buffer */
B200C_API int b200c_carver_bias_add_rgb_area(B200Carver *c, const unsigned char *rgb, int bias_factor, int channels,
                                             int width, int height, int x_off, int y_off);
B200C_API int b200c_carver_rigmask_add_rgb_area(B200Carver *c, const unsigned char *rgb, int channels,
                                                int width, int height, int x_off, int y_off);

/* lqr_carver_build_maps (inside lqr_carver_resize, render.c:318; A.5-A.10): energy map, m-map DP, then the
 * per-seam loop (backtrack, carve, band energy, band DP / side switch) up to `depth`, inflate, width reset.
 * `progress` (may be NULL) is invoked on the calling thread when seam_index % update_step == 0. */
B200C_API int b200c_carver_build_maps(B200Carver *c, int depth, int update_step, b200c_progress_fn progress, void *user);
/* the same with the callback points shifted: `progress` is invoked when (seam_index + update_phase) % update_step == 0.
 * The callback for "i seams done" is delivered only after the device HAS completed them (an event in the queue marks the
 * point; delivery runs one update step behind the enqueue front); a nonzero return stops the session: the seams
 * already queued complete, the call returns B200C_CANCEL and leaves the carver mid-session, as liblqr does. */
B200C_API int b200c_carver_build_maps_phase(B200Carver *c, int depth, int update_step, int update_phase,
                                            b200c_progress_fn progress, void *user);
/* seams of the running (or last) build session the device has completed, read from a mapped host word: never blocks */
B200C_API int b200c_carver_seams_done(const B200Carver *c);
/* The same build_maps session for n independent carvers of equal geometry and knobs (a batch of images, the reference's
 * batch use: batch/batch-gimp-lqr.scm:19-66), advanced in lockstep: ONE launch per step for all of them (image =
 * blockIdx.z), one host thread.  Carvers that do not agree are carved one by one.  Returns when all are done. */
B200C_API int b200c_batch_build_maps(B200Carver **carvers, int n, int depth);
/* lqr_carver_set_width (A.1), mirrored onto attached carvers */
B200C_API int b200c_carver_set_width(B200Carver *c, int w1);
/* lqr_carver_flatten (render.c:325,636; A.11) / transpose (A.11), both recursive on attached carvers */
B200C_API int b200c_carver_flatten(B200Carver *c);
B200C_API int b200c_carver_transpose(B200Carver *c);

enum {
    B200C_W = 0, B200C_H, B200C_W0, B200C_H0, B200C_W_START, B200C_H_START, B200C_LEVEL, B200C_MAX_LEVEL,
    B200C_TRANSPOSED, B200C_CHANNELS, B200C_ACTIVE, B200C_LEFTRIGHT, B200C_DEVICE
};
B200C_API int b200c_carver_get(const B200Carver *c, int field);

/* read-out (lqr_carver_scan_line loop, io_functions.c:155-164; A.12): gathers the visible pixels into an
 * engine-owned pinned host buffer of w*h*channels bytes laid out by internal rows (image rows, or image
 * columns when transposed) and returns it; valid until the next call that changes the carver. */
B200C_API int b200c_carver_readout(B200Carver *c, const unsigned char **host_pixels);
/* the same, for a caller that consumes the rows in order (the shim's scan cursor): large images come back in chunks of
 * rows and the call returns when the first chunk has arrived; b200c_carver_readout_rows(c, row) waits for the chunk
 * that holds internal row `row` and returns how many rows are in the buffer by then (< 0: error), so the copy of the
 * later rows overlaps the caller's handling of the earlier ones (write_carver_to_layer, io_functions.c:155-164). */
B200C_API int b200c_carver_readout_begin(B200Carver *c, const unsigned char **host_pixels);
B200C_API int b200c_carver_readout_rows(B200Carver *c, int row);
/* same gather, left in HBM: d_out must hold w*h*channels bytes (bench.py `value` leg) */
B200C_API int b200c_carver_readout_device(B200Carver *c, void *d_out);
/* lqr_vmap_dump (render.c:725; io_functions.c:216-219; A.13): out = width*height ints in image orientation */
B200C_API int b200c_carver_vmap(B200Carver *c, int *out_host);
/* lqr_carver_get_true_energy: out = w*h floats in image orientation (carver must be in reference state) */
B200C_API int b200c_carver_true_energy(B200Carver *c, float *out_host);

/* carvers created afterwards on this thread queue their work on `stream` (a cudaStream_t; NULL restores the
 * default of one private non-blocking stream per carver).  Lets a host program bracket the engine's kernels
 * with its own CUDA events (bench.py passes torch's current stream). */
B200C_API int b200c_set_stream(void *stream);
/* switches the per-stage CUDA-event timing on/off (same as B200C_TIMING=1 in the environment) */
B200C_API void b200c_set_timing(int on);
/* cumulative count of band cells the incremental DP visited in this process (algorithmic-bytes accounting) */
B200C_API unsigned long long b200c_update_cells(void);

/* waits for all queued device work of this carver */
B200C_API int b200c_carver_sync(B200Carver *c);

/* ---- test / profiling hooks (not used by the shim's API surface) ---------------------------------- */
enum { B200C_DBG_EN = 0, B200C_DBG_M, B200C_DBG_LEAST, B200C_DBG_RAW, B200C_DBG_VS, B200C_DBG_VPATH_X,
       B200C_DBG_BIAS, B200C_DBG_RIGMASK };
/* runs set_width + energy + full DP + `n_seams` iterations of the per-seam loop WITHOUT inflating, so the
 * maps can be compared with the oracle's; the carver is not usable for resize afterwards */
B200C_API int b200c_debug_build(B200Carver *c, int n_seams);
/* copies a device map to host: returns the number of elements written (<= cap), or -1 */
B200C_API long b200c_debug_fetch(B200Carver *c, int what, void *out, long cap);
/* ---- the plug-in's own host loops next to the hot path, device-accelerated.  Host buffers
 * in and out; the calls return when the result is in place.
 *
 * b200c_vmap_colour: the colouring loop of write_vmap_to_layer (reference src/io_functions.c:249-279).  vmap: w*h seam
 * orders (lqr_vmap_get_data), depth = lqr_vmap_get_depth, colours = r,g,b of the GimpRGB pair (render.c:341-342);
 * out_rgba: w*h*4 bytes, byte for byte what the reference loop writes into its rows.
 * b200c_guess_new_size: guess_new_size (reference src/layers_combo.c:274-392) for a discard mask of width x height x
 * bpp placed at (x_off, y_off) over a layer of old_width x old_height; direction 0 = horizontal (new width), 1 =
 * vertical (new height). */
B200C_API int b200c_vmap_colour(const int *vmap, int w, int h, int depth, const double colour_start[3],
                                const double colour_end[3], unsigned char *out_rgba);
B200C_API int b200c_guess_new_size(const unsigned char *mask, int width, int height, int bpp, int has_alpha, int x_off,
                                   int y_off, int old_width, int old_height, int direction, int *new_size);
/* cumulative number of kernel launches issued by this library in this process */
B200C_API long b200c_launch_count(void);
/* cumulative device time (ms) of one stage, measured with CUDA events when B200C_TIMING=1 was set in the
 * environment at load time; stage names: "energy_full","mmap_full","vpath","carve","energy_band",
 * "mmap_update","inflate","readout".  Returns launches counted through *launches. */
B200C_API double b200c_stage_ms(const char *stage, long *launches);
B200C_API void b200c_stage_reset(void);
/* cumulative host wall time (ms, summed over calling threads) the API calls spent in phase `idx`: 0 stream/handle
 * creation, 1 first device allocations, 2 staging-buffer acquire/release, 3 wait for the staging helpers, 4 staging
 * memcpy, 5 H2D enqueue, 6 H2D waits, 7 seam-graph capture + instantiate, 8 seam-graph launches, 9 end-of-loop sync,
 * 10 readout, 11 destroy. */
B200C_API double b200c_hostprof_ms(int idx);

#ifdef __cplusplus
}
#endif
#endif /* B200CARVE_H */
