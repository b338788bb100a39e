// Track/stack compositing node.  Interface parity: lib/toucanRender/Comp.h:11-29.
#pragma once

#include <toucanRender/ImageNode.h>

namespace toucan
{
    //! Compositing node: premultiply the foreground, fit it to the background if the
    //! sizes differ, then Porter-Duff "over".
    class CompNode : public IImageNode
    {
    public:
        CompNode(const std::vector<std::shared_ptr<IImageNode> >& = {});

        virtual ~CompNode();

        //! Set whether images are pre-multiplied before compositing.
        void setPremult(bool);
        bool getPremult() const { return _premult; }

        //! Set whether images are resized before compositing (stored, never consulted:
        //! same as the reference, Comp.cpp:24-27 vs :29-84).
        void setResize(bool);

        Image exec() override;

        //! Composite already evaluated images (used by exec and by the fusion pass).
        Image composite(const Image& fg, const Image& bg) const;

    private:
        bool _premult = false;
        bool _resize = true;
    };
}
