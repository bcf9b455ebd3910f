/* config.h -- what autoconf would generate for the plug-in; only the gettext domain is looked at on the render path. */
#ifndef __CONFIG_STUB_H__
#define __CONFIG_STUB_H__
#define GETTEXT_PACKAGE "gimp20-lqr-plugin"
#define PLUGIN_NAME "gimp-lqr-plugin"
#endif
