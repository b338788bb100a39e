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

        //! Cross-node fusion of same-position operations (tc_pointwise_chain): a node that is such an operation on ONE input
        //! appends its tc_chain_op, names that input and returns true; anything else returns false and ends the chain.
        virtual bool describeChain(tc_chain& chain, std::shared_ptr<IImageNode>& input) const;
        //! The size of the frame exec() will return, when it is known without evaluating anything (false otherwise).
        virtual bool sizeHint(int& width, int& height) const;

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
    //! Per-node device timers (toucan-render -timers, tr_set_node_timers): a Scope brackets the kernels a node enqueues with
    //! two timing events on the calling thread's stream; report() waits for the pending events and returns one line per
    //! label ("label  calls  total ms  mean ms", largest total first).  Off by default: nothing is recorded then.
    class NodeTimers
    {
    public:
        static void setEnabled(bool);
        static bool getEnabled();
        static std::string report(bool reset = true);

        class Scope
        {
        public:
            explicit Scope(const std::string& label);
            ~Scope();
            Scope(const Scope&) = delete;
            Scope& operator=(const Scope&) = delete;

        private:
            std::string _label;
            tc_event _start = nullptr;
        };
    };

    //! Evaluate the chain of same-position nodes that starts at `top` (two or more of them) as ONE kernel over the chain's
    //! source; false when `top` does not start such a chain (nothing has been evaluated then).
    bool tryNodeChain(const IImageNode& top, Image& out);
}
