// Base class of the per-frame image graph.  Interface parity: lib/toucanRender/ImageNode.h:18-57
// -- identical except that exec() returns a device-resident toucan::Image instead of an
// OIIO::ImageBuf (frames stay in HBM across the whole graph).
#pragma once

#include <toucanRender/Image.h>

#include <opentimelineio/otio.h>

#include <memory>
#include <string>
#include <vector>

namespace toucan
{
    class ImageEffectHost;

    //! Base class for image nodes.
    class IImageNode : public std::enable_shared_from_this<IImageNode>
    {
    public:
        IImageNode(
            const std::string& name,
            const std::vector<std::shared_ptr<IImageNode> >& = {});

        virtual ~IImageNode() = 0;

        //! Get the name.
        const std::string& getName() const;

        //! Get the label.
        virtual std::string getLabel() const;

        //! Get the inputs.
        const std::vector<std::shared_ptr<IImageNode> >& getInputs() const;

        //! Set the inputs.
        void setInputs(const std::vector<std::shared_ptr<IImageNode> >&);

        //! Set the time.
        void setTime(const OTIO_NS::RationalTime&);
        const OTIO_NS::RationalTime& getTime() const;

        //! Execute the image operation (enqueues kernels on the thread's stream).
        virtual Image exec() = 0;

        //! Generate a Graphviz graph.
        std::vector<std::string> graph(const std::string& name);

        //! One line per node, children indented: a stable text form of the tree used by the
        //! host-logic tests (no GPU needed).
        virtual std::string describe() const;
        void describeTree(std::vector<std::string>&, int depth = 0) const;

    protected:
        void _graph(const std::shared_ptr<IImageNode>&, std::vector<std::string>&);
        std::string _getGraphName() const;

        std::string _name;
        std::vector<std::shared_ptr<IImageNode> > _inputs;
        OTIO_NS::RationalTime _time;
    };
}
