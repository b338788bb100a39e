// Restatement of the parts of OpenFX 1.5 <ofxCore.h> used on toucan's render path
// (the OpenFX headers are not available in this build environment; names, values and
// struct layouts follow the published OpenFX 1.5 API so third-party hosts/plug-ins stay
// ABI compatible).  Reference use: lib/toucanRender/ImageEffectHost.cpp:22-34,141-211.
#ifndef TOUCAN_OFX_CORE_H
#define TOUCAN_OFX_CORE_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define OfxExport extern __declspec(dllexport)
#else
#define OfxExport extern __attribute__((visibility("default")))
#endif

typedef struct OfxPropertySetStruct* OfxPropertySetHandle;
typedef int OfxStatus;
typedef double OfxTime;

typedef struct OfxHost {
    OfxPropertySetHandle host;
    const void* (*fetchSuite)(OfxPropertySetHandle host, const char* suiteName, int suiteVersion);
} OfxHost;

typedef OfxStatus(OfxPluginEntryPoint)(const char* action, const void* handle, OfxPropertySetHandle inArgs,
                                       OfxPropertySetHandle outArgs);

typedef struct OfxPlugin {
    const char* pluginApi;
    int apiVersion;
    const char* pluginIdentifier;
    unsigned int pluginVersionMajor;
    unsigned int pluginVersionMinor;
    void (*setHost)(OfxHost* host);
    OfxPluginEntryPoint* mainEntry;
} OfxPlugin;

typedef struct OfxRectI { int x1, y1, x2, y2; } OfxRectI;
typedef struct OfxRectD { double x1, y1, x2, y2; } OfxRectD;
typedef struct OfxPointI { int x, y; } OfxPointI;
typedef struct OfxPointD { double x, y; } OfxPointD;
typedef struct OfxRangeD { double min, max; } OfxRangeD;

#define kOfxActionLoad "OfxActionLoad"
#define kOfxActionDescribe "OfxActionDescribe"
#define kOfxActionUnload "OfxActionUnload"
#define kOfxActionCreateInstance "OfxActionCreateInstance"
#define kOfxActionDestroyInstance "OfxActionDestroyInstance"

#define kOfxPropTime "OfxPropTime"
#define kOfxPropLabel "OfxPropLabel"
#define kOfxPropName "OfxPropName"
#define kOfxPropType "OfxPropType"
#define kOfxPropVersion "OfxPropVersion"

#define kOfxStatOK 0
#define kOfxStatFailed ((int)1)
#define kOfxStatErrFatal ((int)2)
#define kOfxStatErrUnknown ((int)3)
#define kOfxStatErrMissingHostFeature ((int)4)
#define kOfxStatErrUnsupported ((int)5)
#define kOfxStatErrExists ((int)6)
#define kOfxStatErrFormat ((int)7)
#define kOfxStatErrMemory ((int)8)
#define kOfxStatErrBadHandle ((int)9)
#define kOfxStatErrBadIndex ((int)10)
#define kOfxStatErrValue ((int)11)
#define kOfxStatReplyYes ((int)12)
#define kOfxStatReplyNo ((int)13)
#define kOfxStatReplyDefault ((int)14)

#define kOfxBitDepthNone "OfxBitDepthNone"
#define kOfxBitDepthByte "OfxBitDepthByte"
#define kOfxBitDepthShort "OfxBitDepthShort"
#define kOfxBitDepthHalf "OfxBitDepthHalf"
#define kOfxBitDepthFloat "OfxBitDepthFloat"

/* the two symbols every plug-in binary exports (lib/toucanRender/PluginUnix.cpp:29-34) */
OfxExport int OfxGetNumberOfPlugins(void);
OfxExport OfxPlugin* OfxGetPlugin(int nth);

#ifdef __cplusplus
}
#endif
#endif
