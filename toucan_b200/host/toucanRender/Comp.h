// Track / stack compositing: the node ImageGraph puts between every video track and what lies below it.
// Same public interface as lib/toucanRender/Comp.h:11-29 (inputs {foreground, background}, setPremult, setResize).
#pragma once

#include <toucanRender/ImageNode.h>

namespace toucan
{
    //! foreground over background: premultiply the foreground (if asked), fit it into the background's frame when
    //! the sizes differ, then Porter-Duff "over".  One fused kernel when the sizes agree.
    class CompNode : public IImageNode
    {
    public:
        CompNode(const ImageNodeList& inputs = {});
        virtual ~CompNode();

        //! Multiply the foreground's colour by its alpha before compositing (ImageGraph always sets this).
        void setPremult(bool);
        bool getPremult() const { return _premultiply; }

        //! Accepted and stored for interface parity; like the reference (Comp.cpp:24-27 against :29-84) nothing
        //! reads it: a foreground of another size is always fitted.
        void setResize(bool);

        Image exec() override;

        //! Over the Fill generator at the bottom of a track stack, with a foreground known to have the Fill's size, the
        //! node is a same-position operation on its foreground (TC_CHAIN_OVER_FILL).
        bool describeChain(tc_chain& chain, std::shared_ptr<IImageNode>& input) const override;
        bool sizeHint(int& width, int& height) const override;

        //! Composite two already evaluated images (exec, and the fusion pass for layers it cannot fuse).
        Image composite(const Image& fg, const Image& bg) const;

    private:
        bool _premultiply = false;
        bool _fit = true;
    };
}
