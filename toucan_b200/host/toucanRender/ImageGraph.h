// The per-frame image graph of a timeline: exec() turns (timeline, time) into a tree of image nodes whose root
// renders the frame.  The public interface is the reference's (lib/toucanRender/ImageGraph.h:21-44: same
// constructor, the three image-info getters, exec with an optional item).  Unlike the reference, exec() keeps its
// working state on its own stack (ImageGraph.cpp: Frame), so only the read-node cache is shared between calls;
// use one ImageGraph per render thread all the same.
#pragma once

#include <toucanRender/ImageNode.h>
#include <toucanRender/TimelineWrapper.h>

#include <ftk/Core/Context.h>

#include <Imath/ImathVec.h>

#include <filesystem>
#include <memory>
#include <string>

namespace toucan
{
    class ImageEffectHost;

    class ImageGraph : public std::enable_shared_from_this<ImageGraph>
    {
    public:
        //! Reads the first video clip that yields an image description (readable media, or a generator with a
        //! "size" parameter) to learn the timeline's frame size, channel count and storage type.
        ImageGraph(
            const std::shared_ptr<ftk::Context>&,
            const std::filesystem::path&,
            const std::shared_ptr<TimelineWrapper>&);
        ~ImageGraph();

        ImageGraph(const ImageGraph&) = delete;
        ImageGraph& operator=(const ImageGraph&) = delete;

        //! Frame size of the timeline in pixels.
        const IMATH_NAMESPACE::V2i& getImageSize() const;
        //! Channels per pixel of the timeline's frames.
        int getImageChannels() const;
        //! Storage type of the timeline's frames: "u8", "u16", "half" or "float".
        const std::string& getImageDataType() const;

        //! The node tree of the frame at `time`.  With `item`, the returned root is that item's output (after its
        //! own effects) instead of the composited stack.
        std::shared_ptr<IImageNode> exec(
            const std::shared_ptr<ImageEffectHost>&,
            const OTIO_NS::RationalTime& time,
            const OTIO_NS::Item* item = nullptr);

    private:
        struct Data;   // timeline facts and the read-node cache
        class Frame;   // builder of one frame's tree
        std::unique_ptr<Data> _d;
    };
}
