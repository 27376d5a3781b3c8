// placeholder, replaced below in this round
#include <VapourSynth4.h>
VS_EXTERNAL_API(void) VapourSynthPluginInit2(VSPlugin *plugin, const VSPLUGINAPI *vspapi) { (void)plugin; (void)vspapi; }
