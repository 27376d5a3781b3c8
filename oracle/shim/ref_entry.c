/* TEST INFRASTRUCTURE (oracle build only).
 * Plugin entry for the oracle build of the reference: registers only the five
 * hot-path filter groups so the out-of-scope translation
 * units (Flow*, Mask, Depan, ...) need not be compiled.  Mirrors what
 * /root/reference/src/EntryPoint.c:28-50 does for those five. */
#include <stdint.h>
#include <stdlib.h>
#include <VapourSynth4.h>
#include "CPU.h"

void mvsuperRegister(VSPlugin *plugin, const VSPLUGINAPI *vspapi);
void mvanalyseRegister(VSPlugin *plugin, const VSPLUGINAPI *vspapi);
void mvdegrainsRegister(VSPlugin *plugin, const VSPLUGINAPI *vspapi);
void mvcompensateRegister(VSPlugin *plugin, const VSPLUGINAPI *vspapi);
void mvrecalculateRegister(VSPlugin *plugin, const VSPLUGINAPI *vspapi);

uint32_t g_cpuinfo = 0;

VS_EXTERNAL_API(void)
VapourSynthPluginInit2(VSPlugin *plugin, const VSPLUGINAPI *vspapi) {
    vspapi->configPlugin("com.nodame.mvtools", "mv", "MVTools v24 (oracle build)", VS_MAKE_VERSION(24, 0), VS_MAKE_VERSION(VAPOURSYNTH_API_MAJOR, VAPOURSYNTH_API_MINOR), 0, plugin);
    mvsuperRegister(plugin, vspapi);
    mvanalyseRegister(plugin, vspapi);
    mvdegrainsRegister(plugin, vspapi);
    mvcompensateRegister(plugin, vspapi);
    mvrecalculateRegister(plugin, vspapi);
    g_cpuinfo = cpu_detect();
}
