// Restatement of OpenFX 1.5 <ofxImageEffect.h> (+ the CUDA render properties of
// <ofxGPURender.h>) limited to what toucan's host and plug-ins use.
// Reference use: lib/toucanRender/ImageEffectHost.cpp:91-97,185-212, ImageEffect.cpp:89-149,
// PropertySet.cpp:431-471, plugins/Util.cpp:10-72.
#ifndef TOUCAN_OFX_IMAGE_EFFECT_H
#define TOUCAN_OFX_IMAGE_EFFECT_H

#include "ofxCore.h"
#include "ofxParam.h"
#include "ofxProperty.h"

#ifdef __cplusplus
extern "C" {
#endif

#define kOfxImageEffectPluginApi "OfxImageEffectPluginAPI"
#define kOfxImageEffectPluginApiVersion 1
#define kOfxImageEffectSuite "OfxImageEffectSuite"

typedef struct OfxImageEffectStruct* OfxImageEffectHandle;
typedef struct OfxImageClipStruct* OfxImageClipHandle;
typedef struct OfxImageMemoryStruct* OfxImageMemoryHandle;

#define kOfxImageEffectActionDescribeInContext "OfxImageEffectActionDescribeInContext"
#define kOfxImageEffectActionRender "OfxImageEffectActionRender"

#define kOfxImageEffectContextGenerator "OfxImageEffectContextGenerator"
#define kOfxImageEffectContextFilter "OfxImageEffectContextFilter"
#define kOfxImageEffectContextTransition "OfxImageEffectContextTransition"
#define kOfxImageEffectContextPaint "OfxImageEffectContextPaint"
#define kOfxImageEffectContextGeneral "OfxImageEffectContextGeneral"
#define kOfxImageEffectContextRetimer "OfxImageEffectContextRetimer"

#define kOfxImageComponentNone "OfxImageComponentNone"
#define kOfxImageComponentRGBA "OfxImageComponentRGBA"
#define kOfxImageComponentRGB "OfxImageComponentRGB"
#define kOfxImageComponentAlpha "OfxImageComponentAlpha"

#define kOfxImageEffectPluginPropGrouping "OfxImageEffectPluginPropGrouping"
#define kOfxImageEffectPropSupportedContexts "OfxImageEffectPropSupportedContexts"
#define kOfxImageEffectPropContext "OfxImageEffectPropContext"
#define kOfxImageEffectPropSupportedComponents "OfxImageEffectPropSupportedComponents"
#define kOfxImageEffectPropSupportedPixelDepths "OfxImageEffectPropSupportedPixelDepths"
#define kOfxImageEffectPropComponents "OfxImageEffectPropComponents"
#define kOfxImageEffectPropPixelDepth "OfxImageEffectPropPixelDepth"
#define kOfxImageEffectPropRenderWindow "OfxImageEffectPropRenderWindow"
#define kOfxImagePropBounds "OfxImagePropBounds"
#define kOfxImagePropRowBytes "OfxImagePropRowBytes"
#define kOfxImagePropData "OfxImagePropData"

/* OpenFX 1.5 CUDA render extension (ofxGPURender.h): with CudaEnabled = 1 in the render
   inArgs, OfxImagePropData of every clip image is a DEVICE pointer and the plug-in must
   enqueue its work on the given stream. */
#define kOfxImageEffectPropCudaRenderSupported "OfxImageEffectPropCudaRenderSupported"
#define kOfxImageEffectPropCudaEnabled "OfxImageEffectPropCudaEnabled"
#define kOfxImageEffectPropCudaStreamSupported "OfxImageEffectPropCudaStreamSupported"
#define kOfxImageEffectPropCudaStream "OfxImageEffectPropCudaStream"


typedef struct OfxImageEffectSuiteV1 {
    OfxStatus (*getPropertySet)(OfxImageEffectHandle imageEffect, OfxPropertySetHandle* propHandle);
    OfxStatus (*getParamSet)(OfxImageEffectHandle imageEffect, OfxParamSetHandle* paramSet);
    OfxStatus (*clipDefine)(OfxImageEffectHandle imageEffect, const char* name, OfxPropertySetHandle* propertySet);
    OfxStatus (*clipGetHandle)(OfxImageEffectHandle imageEffect, const char* name, OfxImageClipHandle* clip, OfxPropertySetHandle* propertySet);
    OfxStatus (*clipGetPropertySet)(OfxImageClipHandle clip, OfxPropertySetHandle* propHandle);
    OfxStatus (*clipGetImage)(OfxImageClipHandle clip, OfxTime time, const OfxRectD* region, OfxPropertySetHandle* imageHandle);
    OfxStatus (*clipReleaseImage)(OfxPropertySetHandle imageHandle);
    OfxStatus (*clipGetRegionOfDefinition)(OfxImageClipHandle clip, OfxTime time, OfxRectD* bounds);
    int (*abort)(OfxImageEffectHandle imageEffect);
    OfxStatus (*imageMemoryAlloc)(OfxImageEffectHandle instanceHandle, size_t nBytes, OfxImageMemoryHandle* memoryHandle);
    OfxStatus (*imageMemoryFree)(OfxImageMemoryHandle memoryHandle);
    OfxStatus (*imageMemoryLock)(OfxImageMemoryHandle memoryHandle, void** returnedPtr);
    OfxStatus (*imageMemoryUnlock)(OfxImageMemoryHandle memoryHandle);
} OfxImageEffectSuiteV1;

#ifdef __cplusplus
}
#endif
#endif
