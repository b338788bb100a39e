#include "ImageNode.h"

#include "StackFusion.h"

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
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

    bool IImageNode::describeChain(tc_chain&, std::shared_ptr<IImageNode>&) const { return false; }

    bool IImageNode::sizeHint(int&, int&) const { return false; }

    namespace
    {
        struct PendingTimer
        {
            std::string label;
            tc_event start, stop;
        };
        std::atomic<bool> timersEnabled{ false };
        std::mutex timersMutex;
        std::vector<PendingTimer> timersPending;
        std::map<std::string, std::pair<unsigned long long, double> > timersTotal; // label -> (calls, ms)
    }

    void NodeTimers::setEnabled(bool value) { timersEnabled = value; }

    bool NodeTimers::getEnabled() { return timersEnabled; }

    NodeTimers::Scope::Scope(const std::string& label)
    {
        if (timersEnabled && DeviceContext::isBound() && TC_OK == tc_event_create_timing(&_start))
        {
            _label = label;
            tc_event_record(_start, DeviceContext::stream());
        }
    }

    NodeTimers::Scope::~Scope()
    {
        if (_start)
        {
            tc_event stop = nullptr;
            if (TC_OK == tc_event_create_timing(&stop))
            {
                tc_event_record(stop, DeviceContext::stream());
                std::lock_guard<std::mutex> lock(timersMutex);
                timersPending.push_back({ _label, _start, stop });
            }
            else
            {
                tc_event_destroy(_start);
            }
        }
    }

    std::string NodeTimers::report(bool reset)
    {
        std::lock_guard<std::mutex> lock(timersMutex);
        for (const auto& p : timersPending)
        {
            float ms = 0.F;
            if (TC_OK == tc_event_sync(p.stop) && TC_OK == tc_event_elapsed_ms(p.start, p.stop, &ms))
            {
                auto& t = timersTotal[p.label];
                t.first += 1;
                t.second += ms;
            }
            tc_event_destroy(p.start);
            tc_event_destroy(p.stop);
        }
        timersPending.clear();
        std::vector<std::pair<std::string, std::pair<unsigned long long, double> > > rows(timersTotal.begin(), timersTotal.end());
        std::sort(rows.begin(), rows.end(), [](const auto& a, const auto& b) { return a.second.second > b.second.second; });
        std::string out;
        for (const auto& r : rows)
        {
            char line[256];
            snprintf(line, sizeof(line), "%-40s %8llu calls %12.3f ms %10.4f ms/call\n", r.first.c_str(), r.second.first, r.second.second,
                     r.second.first ? r.second.second / static_cast<double>(r.second.first) : 0.0);
            out += line;
        }
        if (reset)
        {
            timersTotal.clear();
        }
        return out;
    }

    bool tryNodeChain(const IImageNode& top, Image& out)
    {
        if (!getStackFusionEnabled())
        {
            return false;
        }
        // Walk down while the nodes are same-position operations; ops are recorded top first.
        tc_chain chain;
        memset(&chain, 0, sizeof(chain));
        const IImageNode* node = &top;
        std::shared_ptr<IImageNode> base, next;
        while (node && node->describeChain(chain, next))
        {
            base = next;
            node = base.get();
            const int kind = chain.ops[chain.n - 1].kind;
            if (TC_CHAIN_BLUR == kind || TC_CHAIN_UNSHARP == kind)
            {
                break; // a convolution node can only START a chain: what feeds it is evaluated as its source
            }
        }
        if (chain.n < 2 || !base)
        {
            return false; // a single operation takes the ordinary per-node path
        }
        const Image source = base->exec();
        const ImageSpec& spec = source.spec();
        if (spec.width <= 0 || spec.height <= 0)
        {
            bool comp = false;
            for (int i = 0; i < chain.n; ++i) comp = comp || TC_CHAIN_OVER_FILL == chain.ops[i].kind;
            if (comp)
            {
                return false; // an empty foreground leaves the background: the per-node path renders it (rare: missing media)
            }
            out = Image(); // what a chain of filters over a missing image yields node by node
            return true;
        }
        tc_chain_op ops[TC_CHAIN_MAX];
        for (int i = 0; i < chain.n; ++i)
        {
            ops[i] = chain.ops[chain.n - 1 - i];
            if (TC_CHAIN_OVER_FILL == ops[i].kind && (ops[i].w != spec.width || ops[i].h != spec.height))
            {
                return false; // (sizeHint said otherwise: fall back; the source is evaluated again by the per-node path)
            }
        }
        out = Image(spec);
        const tc_image src = source.view(), dst = out.view();
        NodeTimers::Scope timer("node chain (" + top.getName() + " ...)");
        tcCheck(tc_pointwise_chain(&dst, &src, ops, chain.n, DeviceContext::stream()), "tc_pointwise_chain");
        return true;
    }

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
