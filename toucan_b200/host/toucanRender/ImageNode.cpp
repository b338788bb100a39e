#include "ImageNode.h"

#include <sstream>

namespace toucan
{
    IImageNode::IImageNode(const std::string& name, const ImageNodeList& inputs) :
        _name(name),
        _inputs(inputs)
    {}

    IImageNode::~IImageNode()
    {}

    const std::string& IImageNode::getName() const { return _name; }

    std::string IImageNode::getLabel() const { return _name; }

    const ImageNodeList& IImageNode::getInputs() const { return _inputs; }

    void IImageNode::setInputs(const ImageNodeList& value) { _inputs = value; }

    void IImageNode::setTime(const OTIO_NS::RationalTime& value) { _time = value; }

    const OTIO_NS::RationalTime& IImageNode::getTime() const { return _time; }

    namespace
    {
        // Graphviz identifier of a node: its name made unique by its address
        std::string dotId(const IImageNode& node)
        {
            std::stringstream ss;
            ss << node.getName() << "_" << static_cast<const void*>(&node);
            return ss.str();
        }

        // depth-first: a node's own line, then each input's subtree followed by the edge into the node
        void dotLines(const IImageNode& node, std::vector<std::string>& out)
        {
            out.push_back("    " + dotId(node) + " [label=\"" + node.getLabel() + "\"];");
            for (const auto& input : node.getInputs())
            {
                if (input)
                {
                    dotLines(*input, out);
                    out.push_back("    " + dotId(*input) + " -> " + dotId(node) + ";");
                }
            }
        }
    }

    std::vector<std::string> IImageNode::graph(const std::string& name)
    {
        std::vector<std::string> out;
        out.push_back("digraph " + name + " {");
        out.push_back("    node [shape=box, fontsize=12, margin=0.05, width=0, height=0];");
        dotLines(*this, out);
        out.push_back("}");
        return out;
    }

    std::string IImageNode::describe() const { return getLabel(); }

    void IImageNode::describeTree(std::vector<std::string>& out, int depth) const
    {
        out.push_back(std::string(static_cast<size_t>(depth) * 2, ' ') + describe());
        for (const auto& input : _inputs)
        {
            if (input)
            {
                input->describeTree(out, depth + 1);
            }
            else
            {
                out.push_back(std::string(static_cast<size_t>(depth + 1) * 2, ' ') + "(null)");
            }
        }
    }
}
