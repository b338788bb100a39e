// Restatement of OpenFX 1.5 <ofxParam.h>: parameter suite layout and the names used by
// toucan's plug-ins.  The reference host serves only paramDefine, paramGetHandle and
// paramGetValue (lib/toucanRender/ImageEffectHost.cpp:86-89); this host fills the same
// three slots and leaves the others null.
#ifndef TOUCAN_OFX_PARAM_H
#define TOUCAN_OFX_PARAM_H

#include "ofxCore.h"

#ifdef __cplusplus
extern "C" {
#endif

#define kOfxParameterSuite "OfxParameterSuite"

typedef struct OfxParamStruct* OfxParamHandle;
typedef struct OfxParamSetStruct* OfxParamSetHandle;

#define kOfxParamTypeInteger "OfxParamTypeInteger"
#define kOfxParamTypeDouble "OfxParamTypeDouble"
#define kOfxParamTypeBoolean "OfxParamTypeBoolean"
#define kOfxParamTypeChoice "OfxParamTypeChoice"
#define kOfxParamTypeRGBA "OfxParamTypeRGBA"
#define kOfxParamTypeRGB "OfxParamTypeRGB"
#define kOfxParamTypeDouble2D "OfxParamTypeDouble2D"
#define kOfxParamTypeInteger2D "OfxParamTypeInteger2D"
#define kOfxParamTypeDouble3D "OfxParamTypeDouble3D"
#define kOfxParamTypeInteger3D "OfxParamTypeInteger3D"
#define kOfxParamTypeString "OfxParamTypeString"
#define kOfxParamTypeCustom "OfxParamTypeCustom"
#define kOfxParamTypeGroup "OfxParamTypeGroup"
#define kOfxParamTypePage "OfxParamTypePage"
#define kOfxParamTypePushButton "OfxParamTypePushButton"

#define kOfxParamPropDefault "OfxParamPropDefault"
#define kOfxParamPropHint "OfxParamPropHint"
#define kOfxParamPropMin "OfxParamPropMin"
#define kOfxParamPropMax "OfxParamPropMax"

typedef struct OfxParameterSuiteV1 {
    OfxStatus (*paramDefine)(OfxParamSetHandle paramSet, const char* paramType, const char* name, OfxPropertySetHandle* propertySet);
    OfxStatus (*paramGetHandle)(OfxParamSetHandle paramSet, const char* name, OfxParamHandle* param, OfxPropertySetHandle* propertySet);
    OfxStatus (*paramSetGetPropertySet)(OfxParamSetHandle paramSet, OfxPropertySetHandle* propHandle);
    OfxStatus (*paramGetPropertySet)(OfxParamHandle param, OfxPropertySetHandle* propHandle);
    OfxStatus (*paramGetValue)(OfxParamHandle paramHandle, ...);
    OfxStatus (*paramGetValueAtTime)(OfxParamHandle paramHandle, OfxTime time, ...);
    OfxStatus (*paramGetDerivative)(OfxParamHandle paramHandle, OfxTime time, ...);
    OfxStatus (*paramGetIntegral)(OfxParamHandle paramHandle, OfxTime time1, OfxTime time2, ...);
    OfxStatus (*paramSetValue)(OfxParamHandle paramHandle, ...);
    OfxStatus (*paramSetValueAtTime)(OfxParamHandle paramHandle, OfxTime time, ...);
    OfxStatus (*paramGetNumKeys)(OfxParamHandle paramHandle, unsigned int* numberOfKeys);
    OfxStatus (*paramGetKeyTime)(OfxParamHandle paramHandle, unsigned int nthKey, OfxTime* time);
    OfxStatus (*paramGetKeyIndex)(OfxParamHandle paramHandle, OfxTime time, int direction, int* index);
    OfxStatus (*paramDeleteKey)(OfxParamHandle paramHandle, OfxTime time);
    OfxStatus (*paramDeleteAllKeys)(OfxParamHandle paramHandle);
    OfxStatus (*paramCopy)(OfxParamHandle paramTo, OfxParamHandle paramFrom, OfxTime dstOffset, const OfxRangeD* frameRange);
    OfxStatus (*paramEditBegin)(OfxParamSetHandle paramSet, const char* name);
    OfxStatus (*paramEditEnd)(OfxParamSetHandle paramSet);
} OfxParameterSuiteV1;

#ifdef __cplusplus
}
#endif
#endif
