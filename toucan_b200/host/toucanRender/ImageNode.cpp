#include "ImageNode.h"

#include <sstream>

namespace toucan
{
    IImageNode::IImageNode(
        const std::string& name,
        const std::vector<std::shared_ptr<IImageNode> >& inputs) :
        _name(name),
        _inputs(inputs)
    {}

    IImageNode::~IImageNode()
    {}

    const std::string& IImageNode::getName() const { return _name; }

    std::string IImageNode::getLabel() const { return _name; }

    const std::vector<std::shared_ptr<IImageNode> >& IImageNode::getInputs() const { return _inputs; }

    void IImageNode::setInputs(const std::vector<std::shared_ptr<IImageNode> >& value) { _inputs = value; }

    void IImageNode::setTime(const OTIO_NS::RationalTime& value) { _time = value; }

    const OTIO_NS::RationalTime& IImageNode::getTime() const { return _time; }

    std::vector<std::string> IImageNode::graph(const std::string& name)
    {
        std::vector<std::string> out;
        out.push_back("digraph " + name + " {");
        out.push_back("    node [shape=box, fontsize=12, margin=0.05, width=0, height=0];");
        _graph(shared_from_this(), out);
        out.push_back("}");
        return out;
    }

    void IImageNode::_graph(const std::shared_ptr<IImageNode>& node, std::vector<std::string>& out)
    {
        out.push_back("    " + node->_getGraphName() + " [label=\"" + node->getLabel() + "\"];");
        for (const auto& input : node->_inputs)
        {
            if (input)
            {
                _graph(input, out);
                out.push_back("    " + input->_getGraphName() + " -> " + node->_getGraphName() + ";");
            }
        }
    }

    std::string IImageNode::_getGraphName() const
    {
        std::stringstream ss;
        ss << _name << "_" << static_cast<const void*>(this);
        return ss.str();
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
