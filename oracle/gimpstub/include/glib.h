/* glib.h -- the handful of glib names the reference's render path uses (src/render.c, src/io_functions.c), so those
 * files compile UNMODIFIED in a container without glib.  Test infrastructure (oracle/_ref), not product code. */
#ifndef __G_LIB_STUB_H__
#define __G_LIB_STUB_H__
#define __G_TYPES_H__
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef int gint;
typedef unsigned int guint;
typedef unsigned char guchar;
typedef char gchar;
typedef float gfloat;
typedef double gdouble;
typedef int gboolean;
typedef void *gpointer;
typedef int gint32;
typedef unsigned long gsize;

#ifndef TRUE
#define TRUE 1
#endif
#ifndef FALSE
#define FALSE 0
#endif
#ifndef MAX
#define MAX(a, b) (((a) > (b)) ? (a) : (b))
#endif
#ifndef MIN
#define MIN(a, b) (((a) < (b)) ? (a) : (b))
#endif
#define G_STMT_START do
#define G_STMT_END while (0)

#define g_try_new(type, n) ((type *) malloc(sizeof(type) * (size_t) ((n) > 0 ? (n) : 1)))
#define g_new(type, n) ((type *) malloc(sizeof(type) * (size_t) ((n) > 0 ? (n) : 1)))
#define g_free(p) free(p)
#define g_snprintf snprintf
void g_message(const gchar *format, ...);
#endif
