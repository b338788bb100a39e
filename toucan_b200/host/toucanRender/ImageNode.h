// A node of the per-frame image graph.  The public interface is the reference's IImageNode
// (lib/toucanRender/ImageNode.h:18-57: name, label, inputs, time, exec, graph) with one change of type:
// exec() returns a device-resident toucan::Image instead of an OIIO::ImageBuf, so frames stay in HBM
// across the whole graph and only the writer downloads.
#pragma once

#include <toucanRender/Image.h>

#include <opentimelineio/otio.h>

#include <memory>
#include <string>
#include <vector>

namespace toucan
{
    class ImageEffectHost;
    class IImageNode;

    using ImageNodeList = std::vector<std::shared_ptr<IImageNode> >;

    class IImageNode : public std::enable_shared_from_this<IImageNode>
    {
    public:
        IImageNode(const std::string& name, const ImageNodeList& inputs = {});
        virtual ~IImageNode() = 0;

        // ---- identity and wiring ----
        const std::string& getName() const;
        //! What graph() and describe() print for this node; the name unless a subclass adds detail.
        virtual std::string getLabel() const;
        const ImageNodeList& getInputs() const;
        void setInputs(const ImageNodeList&);

        // ---- the time the node is evaluated at (read nodes pick their frame from it) ----
        void setTime(const OTIO_NS::RationalTime&);
        const OTIO_NS::RationalTime& getTime() const;

        //! Evaluate the inputs, then this node: enqueues kernels on the calling thread's stream and returns the
        //! node's output without waiting for them.
        virtual Image exec() = 0;

        //! The tree under this node as Graphviz source, one line per element.
        std::vector<std::string> graph(const std::string& name);

        //! One line per node, children indented: a stable text form of the tree used by the
        //! host-logic tests (no GPU needed).
        virtual std::string describe() const;
        void describeTree(std::vector<std::string>&, int depth = 0) const;

    protected:
        std::string _name;
        ImageNodeList _inputs;
        OTIO_NS::RationalTime _time;
    };
}
