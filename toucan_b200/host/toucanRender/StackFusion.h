// Cross-node fusion of the multi-track hot chain.
//
// ImageGraph::exec builds, for a timeline of N tracks, the chain
//     Comp(track_N, Comp(track_N-1, ... Comp(track_1, Fill(background))))
// (lib/toucanRender/ImageGraph.cpp:165-198) which the reference evaluates node by node, one
// full-frame buffer per node.  When every track_i is  [Dissolve of]  [Blur of]  a float32
// source of the timeline's size, the whole chain is one launch of tc_stack_f32: every source
// is read once, the composite is written once, and no intermediate touches HBM.
#pragma once

#include <toucanRender/Image.h>
#include <toucanRender/ImageNode.h>

#include <Imath/ImathVec.h>

namespace toucan
{
    class CompNode;

    //! Evaluate the chain rooted at `top` in one pass if it matches; false = not fusable
    //! (the caller then evaluates node by node).
    bool tryFusedStack(const CompNode& top, Image& out);

    //! Is `node` the Fill generator a track stack stands on?  Yields its size and the colour its render body would
    //! receive (unquantised).  Used to composite over a constant instead of materialising the background frame.
    bool matchFillBackground(const std::shared_ptr<IImageNode>& node, IMATH_NAMESPACE::V2i& size, float color[4]);

    //! Enable / disable the fusion pass for the calling process (default: enabled; the
    //! environment variable TOUCAN_NO_FUSION=1 disables it at start-up).
    void setStackFusionEnabled(bool);
    bool getStackFusionEnabled();
}
