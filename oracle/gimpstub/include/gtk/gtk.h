/* gtk/gtk.h -- src/render.c includes it but uses nothing from it. */
#ifndef __GTK_STUB_H__
#define __GTK_STUB_H__
#include <glib.h>
typedef struct _GtkWidget GtkWidget;
#endif
