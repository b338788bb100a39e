// Table-driven OpenFX image-effect modules for the toucan render path.
//
// The reference implements each of its 21 effects as a hand-written class with its own
// describe / createInstance / render overrides and a per-instance handle cache
// (plugins/Plugin.h:12-47, plugins/FilterPlugin.h:12-330, ...).  Here an effect is one row
// of a table -- identifier, context, described parameters, render body -- and ONE engine
// serves the OpenFX actions for every row.  The render bodies call the sm_100a kernels
// through the C ABI (include/toucan_cuda.h) on the device pointers and CUDA stream the
// host passes with the OpenFX CUDA-render properties.
//
// Thread safety: a module keeps no per-instance state (parameters are fetched from the
// host inside Render), so CreateInstance / Render / DestroyInstance may run concurrently
// from one render thread per GPU.  The reference's per-instance std::map caches
// (plugins/FilterPlugin.h:64) are not thread safe.
#pragma once

#include <OpenFX/ofxImageEffect.h>
#include <OpenFX/ofxParam.h>

#include <toucan_cuda.h>

#include <cstdint>
#include <string>
#include <vector>

namespace toucan_ofx
{
    enum class Context { Generator, Filter, Transition };

    //! A described parameter (OfxParamPropDefault as the reference describes it).
    struct ParamDef
    {
        enum Kind { Integer, Integer2D, Boolean, Double, String, RGBA };
        const char* name;
        Kind kind;
        const char* label;
        double d[4];     // Double: d[0]; RGBA: d[0..3]
        int i[2];        // Integer / Boolean: i[0]; Integer2D: i[0..1]
        const char* s;   // String
    };

    //! Typed access to the parameters of the instance being rendered.  Every getter takes
    //! the plug-in's LOCAL initialiser: it is what the effect uses when the host does not
    //! deliver the parameter (described int/bool defaults never arrive).
    class Params
    {
    public:
        Params(const OfxParameterSuiteV1&, OfxParamSetHandle);
        double getDouble(const char* name, double local) const;
        int64_t getInt(const char* name, int64_t local) const;
        //! Reads a Boolean into a bool (Gradient's `vertical`) ...
        bool getBool(const char* name, bool local) const;
        //! ... or into the low byte of an int (Noise's `mono`, ColorConvert's `premult`).
        int getBoolAsInt(const char* name, int local) const;
        std::string getString(const char* name, const std::string& local) const;
        void getInt2(const char* name, int64_t out[2]) const;
        void getRGBA(const char* name, double out[4]) const;

    private:
        OfxParamHandle _handle(const char* name) const;
        const OfxParameterSuiteV1& _suite;
        OfxParamSetHandle _set;
    };

    //! What a render body receives: device images + the stream to enqueue on.
    struct RenderArgs
    {
        tc_image output;
        tc_image source[2]; // Filter: [0] = Source; Transition: [0] = SourceFrom, [1] = SourceTo
        OfxTime time;
        OfxRectI window;
        tc_stream stream;
        //! Non-null when the host asked the effect to DESCRIBE its pixel operation instead of launching it
        //! (render-argument property kToucanPropPointwiseChain): the body appends one tc_chain_op and returns.
        //! Only effects flagged `pointwise` are ever called this way; they have no images then.
        tc_chain* chain;
    };

    //! Render-argument property (pointer to a tc_chain): the host collects consecutive pointwise effects of a
    //! graph into one kernel launch.  An effect that is not pointwise answers kOfxStatReplyDefault.
    constexpr const char* kToucanPropPointwiseChain = "ToucanPropPointwiseChain";

    //! Append `op` to the chain being recorded.
    int chainAppend(const RenderArgs&, const tc_chain_op&);

    inline tc_chain_op chainOp(int kind, float value = 0.F)
    {
        tc_chain_op op = {};
        op.kind = kind;
        op.value = value;
        return op;
    }

    using RenderFn = int (*)(const RenderArgs&, const Params&); // returns a tc_status

    struct EffectDef
    {
        const char* identifier; // "toucan:Blur"
        const char* label;      // "Blur"
        Context context;
        std::vector<ParamDef> params;
        RenderFn render;
        bool pointwise = false; //!< the render body honours RenderArgs::chain
    };

    //! The effect table of this module (defined once per .ofx).
    const std::vector<EffectDef>& effects();

    inline ParamDef paramDouble(const char* n, const char* l, double v) { return { n, ParamDef::Double, l, { v, 0, 0, 0 }, { 0, 0 }, "" }; }
    inline ParamDef paramString(const char* n, const char* l, const char* v) { return { n, ParamDef::String, l, { 0, 0, 0, 0 }, { 0, 0 }, v }; }
    inline ParamDef paramInt(const char* n, const char* l, int v) { return { n, ParamDef::Integer, l, { 0, 0, 0, 0 }, { v, 0 }, "" }; }
    inline ParamDef paramBool(const char* n, const char* l, int v) { return { n, ParamDef::Boolean, l, { 0, 0, 0, 0 }, { v, 0 }, "" }; }
    inline ParamDef paramInt2(const char* n, const char* l, int a, int b) { return { n, ParamDef::Integer2D, l, { 0, 0, 0, 0 }, { a, b }, "" }; }
    inline ParamDef paramRGBA(const char* n, const char* l, double r, double g, double b, double a) { return { n, ParamDef::RGBA, l, { r, g, b, a }, { 0, 0 }, "" }; }
}
