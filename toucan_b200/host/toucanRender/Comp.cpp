#include "Comp.h"

#include "StackFusion.h"
#include "Util.h"

#include <cstring>

namespace toucan
{
    CompNode::CompNode(const std::vector<std::shared_ptr<IImageNode> >& inputs) :
        IImageNode("Comp", inputs)
    {}

    CompNode::~CompNode()
    {}

    void CompNode::setPremult(bool premult) { _premultiply = premult; }

    void CompNode::setResize(bool resize) { _fit = resize; }

    Image CompNode::composite(const Image& fgIn, const Image& bg) const
    {
        const ImageSpec& fgSpec = fgIn.spec();
        const ImageSpec& bgSpec = bg.spec();
        const bool fgValid = fgSpec.width > 0 && fgSpec.height > 0;
        const bool bgValid = bgSpec.width > 0 && bgSpec.height > 0;
        if (!fgValid)
        {
            return bg;
        }
        tc_stream stream = DeviceContext::stream();
        if (bgValid && (fgSpec.width != bgSpec.width || fgSpec.height != bgSpec.height))
        {
            // Slow path (Comp.cpp:37-68): premult -> resize into the fitted box -> paste on a
            // transparent buffer of the background's size and type -> over.
            Image fg = fgIn;
            if (_premultiply)
            {
                Image pm(fgSpec);
                const tc_image s = fgIn.view(), d = pm.view();
                tcCheck(tc_premult(&d, &s, stream), "tc_premult");
                fg = pm;
            }
            const IMATH_NAMESPACE::Box2i box = fit(
                IMATH_NAMESPACE::V2i(bgSpec.width, bgSpec.height),
                IMATH_NAMESPACE::V2i(fgSpec.width, fgSpec.height));
            const int bw = box.max.x - box.min.x + 1;
            const int bh = box.max.y - box.min.y + 1;
            Image pasted(ImageSpec(bgSpec.width, bgSpec.height, bgSpec.nchannels, bgSpec.format));
            if (bw > 0 && bh > 0)
            {
                Image resized(ImageSpec(bw, bh, fgSpec.nchannels, fgSpec.format));
                const tc_image s = fg.view(), r = resized.view(), p = pasted.view();
                tcCheck(tc_resize(&r, &s, "", 0.f, stream), "tc_resize");
                tcCheck(tc_paste_on_black(&p, box.min.x, box.min.y, &r, stream), "tc_paste_on_black");
            }
            else
            {
                tcCheck(tc_memset(pasted.data(), 0, pasted.spec().image_bytes(), stream), "tc_memset");
            }
            Image out(ImageSpec(bgSpec.width, bgSpec.height, 4, tc_type_merge(bgSpec.format, bgSpec.format)));
            const tc_image f = pasted.view(), b = bg.view(), o = out.view();
            tcCheck(tc_over(&o, &f, &b, 0, stream), "tc_over");
            return out;
        }
        if (!bgValid)
        {
            // over() with an empty background: the foreground alone
            if (!_premultiply) return fgIn;
            Image pm(fgSpec);
            const tc_image s = fgIn.view(), d = pm.view();
            tcCheck(tc_premult(&d, &s, stream), "tc_premult");
            return pm;
        }
        // Fast path: premult + over fused into one pass.
        Image out(ImageSpec(bgSpec.width, bgSpec.height, 4, tc_type_merge(fgSpec.format, bgSpec.format)));
        const tc_image f = fgIn.view(), b = bg.view(), o = out.view();
        tcCheck(tc_over(&o, &f, &b, _premultiply ? 1 : 0, stream), "tc_over");
        return out;
    }

    bool CompNode::describeChain(tc_chain& chain, std::shared_ptr<IImageNode>& input) const
    {
        if (_inputs.size() < 2 || !_inputs[0] || !_inputs[1] || chain.n >= TC_CHAIN_MAX)
        {
            return false;
        }
        IMATH_NAMESPACE::V2i fillSize(0, 0);
        float fillColor[4];
        int w = 0, h = 0;
        if (!matchFillBackground(_inputs[1], fillSize, fillColor) || !_inputs[0]->sizeHint(w, h) || w != fillSize.x || h != fillSize.y)
        {
            return false;
        }
        tc_chain_op& op = chain.ops[chain.n++];
        memset(&op, 0, sizeof(op));
        op.kind = TC_CHAIN_OVER_FILL;
        op.flag = _premultiply ? 1 : 0;
        for (int c = 0; c < 4; ++c) op.color[c] = fillColor[c];
        op.w = fillSize.x;
        op.h = fillSize.y;
        input = _inputs[0];
        return true;
    }

    bool CompNode::sizeHint(int& width, int& height) const
    {
        // the composite has the background's size (a foreground of another size is fitted into it)
        if (_inputs.size() > 1 && _inputs[1]) return _inputs[1]->sizeHint(width, height);
        if (1 == _inputs.size() && _inputs[0]) return _inputs[0]->sizeHint(width, height);
        return false;
    }

    Image CompNode::exec()
    {
        Image buf;
        if (_inputs.size() > 1 && _inputs[0] && _inputs[1])
        {
            // Cross-node fusion: a chain of CompNodes over [Dissolve of] [Blur of] float
            // sources collapses into one pass (StackFusion.h).  Falls through when the
            // pattern does not match.
            if (_premultiply && tryFusedStack(*this, buf))
            {
                return buf;
            }
            // ... and over the Fill's colour the node is one more step of the foreground's chain of same-position operations
            if (tryNodeChain(*this, buf))
            {
                return buf;
            }
            const Image fg = _inputs[0]->exec();
            // The bottom of a track stack is a Fill generator (ImageGraph.cpp:165-171): composite over its colour
            // instead of rendering, writing and re-reading a constant frame.
            IMATH_NAMESPACE::V2i fillSize(0, 0);
            float fillColor[4];
            const ImageSpec& fgSpec = fg.spec();
            if (getStackFusionEnabled() && fgSpec.width > 0 && fgSpec.height > 0 && matchFillBackground(_inputs[1], fillSize, fillColor) &&
                fillSize.x == fgSpec.width && fillSize.y == fgSpec.height)
            {
                // generators render 4-channel UINT8 (ImageEffect.cpp:93)
                buf = Image(ImageSpec(fgSpec.width, fgSpec.height, 4, tc_type_merge(fgSpec.format, TC_U8)));
                const tc_image f = fg.view(), o = buf.view();
                NodeTimers::Scope timer("Comp (over the Fill colour)");
                tcCheck(tc_over_fill(&o, &f, fillColor, TC_U8, _premultiply ? 1 : 0, DeviceContext::stream()), "tc_over_fill");
                return buf;
            }
            const Image bg = _inputs[1]->exec();
            NodeTimers::Scope timer("Comp");
            buf = composite(fg, bg);
        }
        else if (1 == _inputs.size() && _inputs[0])
        {
            buf = _inputs[0]->exec();
            if (_premultiply && buf.spec().width > 0 && buf.spec().height > 0)
            {
                Image pm(buf.spec());
                const tc_image s = buf.view(), d = pm.view();
                tcCheck(tc_premult(&d, &s, DeviceContext::stream()), "tc_premult");
                buf = pm;
            }
        }
        return buf;
    }
}
