// plugin_internal.h — what the two plug-in translation units (plugin.cu: simple codecs, engine pool, combiners; plugin_complex.cu:
// DOMQ / ACGT / PBWT / LONGR) share.
#pragma once
#include "../../include/gzb200.h"

namespace gzbp {

extern gzb_plugin_host  g_host;
extern gzb_plugin_host2 g_host2;
extern int g_n_devices;

[[noreturn]] void plugin_abort (const char *what, const char *name, const char *detail);

// An engine checked out of the process-wide pool for the duration of one plug-in call.  genozip starts a new pthread for every
// VBlock (src/dispatcher.c:335), so an engine cannot belong to a thread: it belongs to the pool, per device, and is reused.
struct EngineLease {
    gzb_engine *e; int dev;
    EngineLease (VBlockP vb, const char *name);
    ~EngineLease ();
    EngineLease (const EngineLease &) = delete;
};

// one section through the simple codecs: directly, or through the process-wide combiners when combining is on
int run_section (gzb_engine *e, int dev, gzb_section *s, bool compress);

} // namespace gzbp
