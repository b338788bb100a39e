#include "ImageGraph.h"

#include "Comp.h"
#include "ImageEffectHost.h"
#include "Read.h"
#include "Util.h"

namespace toucan
{
    namespace
    {
        const std::string logPrefix = "toucan::ImageGraph";
    }

    // What the constructor learns about the timeline, and the read-node cache shared by every frame.
    struct ImageGraph::Data
    {
        std::weak_ptr<ftk::Context> context;
        std::filesystem::path path;
        std::shared_ptr<TimelineWrapper> timelineWrapper;
        OTIO_NS::TimeRange timeRange;
        IMATH_NAMESPACE::V2i imageSize = IMATH_NAMESPACE::V2i(0, 0);
        int imageChannels = 0;
        std::string imageDataType;
        ftk::LRUCache<const OTIO_NS::MediaReference*, std::shared_ptr<IReadNode> > readCache;

        std::shared_ptr<IReadNode> read(const OTIO_NS::MediaReference*, bool cache);
        void imageInfo(const std::shared_ptr<IReadNode>&);
    };

    // Builder of ONE frame's node tree: the state the reference keeps in ImageGraph members for the duration of
    // exec() (ImageGraph.h:74-77) lives here, on exec()'s stack.
    class ImageGraph::Frame
    {
    public:
        using EffectList = std::vector<OTIO_NS::SerializableObject::Retainer<OTIO_NS::Effect> >;

        Frame(Data& data, const std::shared_ptr<ImageEffectHost>& h, const OTIO_NS::Item* solo) : d(data), host(h), soloItem(solo) {}

        std::shared_ptr<IImageNode> buildTrack(
            const OTIO_NS::RationalTime&,
            const OTIO_NS::SerializableObject::Retainer<OTIO_NS::Track>&);
        std::shared_ptr<IImageNode> buildItem(
            const OTIO_NS::RationalTime&,
            const OTIO_NS::SerializableObject::Retainer<OTIO_NS::Item>&);
        std::shared_ptr<IImageNode> buildTransition(
            const OTIO_NS::SerializableObject::Retainer<OTIO_NS::Transition>&,
            const OTIO_NS::TimeRange&,
            const OTIO_NS::RationalTime&,
            const std::shared_ptr<IImageNode>& from,
            const std::shared_ptr<IImageNode>& to);
        OTIO_NS::RationalTime warpTime(const OTIO_NS::RationalTime&, const OTIO_NS::TimeRange&, const EffectList&);
        std::shared_ptr<IImageNode> applyEffects(const OTIO_NS::RationalTime&, const EffectList&, const std::shared_ptr<IImageNode>&);

        Data& d;
        std::shared_ptr<ImageEffectHost> host;
        const OTIO_NS::Item* soloItem = nullptr;   // exec()'s optional item: its output replaces the stack's
        std::shared_ptr<IImageNode> soloNode;
    };

    ImageGraph::ImageGraph(
        const std::shared_ptr<ftk::Context>& context,
        const std::filesystem::path& path,
        const std::shared_ptr<TimelineWrapper>& timelineWrapper) :
        _d(new Data)
    {
        Data& d = *_d;
        d.context = context;
        d.path = path;
        d.timelineWrapper = timelineWrapper;
        d.timeRange = timelineWrapper->getTimeRange();
        d.readCache.setMax(20);

        // The timeline's image size / channels / type come from the first video clip that
        // yields them: readable media, or a generator with a "size" parameter (generators
        // are hard-wired to 4 x UINT8).  ImageGraph.cpp:53-123.
        for (const auto& clip : getVideoClips(d.timelineWrapper->getTimeline()))
        {
            OTIO_NS::MediaReference* ref = clip->media_reference();
            if (auto generatorRef = dynamic_cast<OTIO_NS::GeneratorReference*>(ref))
            {
                const auto& parameters = generatorRef->parameters();
                const auto i = parameters.find("size");
                if (i != parameters.end() && i->second.has_value())
                {
                    anyToVec(std::any_cast<OTIO_NS::AnyVector>(i->second), d.imageSize);
                    d.imageChannels = 4;
                    d.imageDataType = toImageDataType(TC_U8);
                    break;
                }
            }
            else if (auto read = d.read(ref, false))
            {
                if (read->getSpec().width > 0)
                {
                    d.imageInfo(read);
                    break;
                }
            }
        }
    }

    ImageGraph::~ImageGraph()
    {}

    const IMATH_NAMESPACE::V2i& ImageGraph::getImageSize() const { return _d->imageSize; }

    int ImageGraph::getImageChannels() const { return _d->imageChannels; }

    const std::string& ImageGraph::getImageDataType() const { return _d->imageDataType; }

    void ImageGraph::Data::imageInfo(const std::shared_ptr<IReadNode>& read)
    {
        const ImageSpec& spec = read->getSpec();
        imageSize.x = spec.width;
        imageSize.y = spec.height;
        imageChannels = spec.nchannels;
        imageDataType = toImageDataType(spec.format);
    }

    std::shared_ptr<IReadNode> ImageGraph::Data::read(const OTIO_NS::MediaReference* ref, bool cache)
    {
        // External and image-sequence references become read nodes; a reader that cannot be
        // opened is logged and yields no node (the clip then contributes nothing).
        std::shared_ptr<IReadNode> read;
        if (!dynamic_cast<const OTIO_NS::ExternalReference*>(ref) &&
            !dynamic_cast<const OTIO_NS::ImageSequenceReference*>(ref))
        {
            return read;
        }
        if (cache && readCache.get(ref, read))
        {
            return read;
        }
        try
        {
            read = timelineWrapper->createReadNode(ref);
            if (cache)
            {
                readCache.add(ref, read);
            }
        }
        catch (const std::exception& e)
        {
            if (auto ctx = context.lock())
            {
                ctx->getSystem<ftk::LogSystem>()->print(logPrefix, e.what(), ftk::LogType::Error);
            }
        }
        return read;
    }

    std::shared_ptr<IImageNode> ImageGraph::exec(
        const std::shared_ptr<ImageEffectHost>& host,
        const OTIO_NS::RationalTime& time,
        const OTIO_NS::Item* itemNode)
    {
        Data& d = *_d;
        Frame frame(d, host, itemNode);

        // Background: opaque black at the timeline's size.
        OTIO_NS::AnyDictionary metaData;
        metaData["size"] = vecToAny(d.imageSize);
        metaData["color"] = vecToAny(IMATH_NAMESPACE::V4f(0.F, 0.F, 0.F, 1.F));
        std::shared_ptr<IImageNode> node = host->createNode(metaData, "toucan:Fill");

        auto stack = d.timelineWrapper->getTimeline()->tracks();
        const Frame::EffectList& stackEffects = stack->effects();
        OTIO_NS::RationalTime t = time - d.timeRange.start_time();
        t = frame.warpTime(t, stack->available_range(), stackEffects);

        // Bottom track first; each video track with clips is composited over what is below.
        for (const auto& child : stack->children())
        {
            auto track = OTIO_NS::dynamic_retainer_cast<OTIO_NS::Track>(child);
            if (!track || track->kind() != OTIO_NS::Track::Kind::video || track->find_clips().empty())
            {
                continue;
            }
            const Frame::EffectList& trackEffects = track->effects();
            OTIO_NS::RationalTime t2 = t;
            if (!trackEffects.empty())
            {
                t2 = frame.warpTime(t2, track->available_range(), trackEffects);
            }
            auto trackNode = frame.applyEffects(t2, trackEffects, frame.buildTrack(t2, track));

            std::vector<std::shared_ptr<IImageNode> > nodes;
            if (trackNode) nodes.push_back(trackNode);
            if (node) nodes.push_back(node);
            auto comp = std::make_shared<CompNode>(nodes);
            comp->setPremult(true);
            node = comp;
        }

        node = frame.applyEffects(t, stackEffects, node);

        // a caller that asked for one item's output gets that node instead of the stack's
        return frame.soloNode ? frame.soloNode : node;
    }

    std::shared_ptr<IImageNode> ImageGraph::Frame::buildTransition(
        const OTIO_NS::SerializableObject::Retainer<OTIO_NS::Transition>& transition,
        const OTIO_NS::TimeRange& range,
        const OTIO_NS::RationalTime& time,
        const std::shared_ptr<IImageNode>& from,
        const std::shared_ptr<IImageNode>& to)
    {
        // progress through the transition, injected as the "value" parameter
        const double value = (time - range.start_time()).value() / range.duration().value();
        OTIO_NS::AnyDictionary metaData = transition->metadata();
        metaData["value"] = value;
        auto node = host->createNode(metaData, transition->transition_type(), { from, to });
        if (!node)
        {
            // unknown transition types fall back to a dissolve
            node = host->createNode(metaData, "toucan:Dissolve", { from, to });
        }
        return node;
    }

    std::shared_ptr<IImageNode> ImageGraph::Frame::buildTrack(
        const OTIO_NS::RationalTime& time,
        const OTIO_NS::SerializableObject::Retainer<OTIO_NS::Track>& track)
    {
        std::shared_ptr<IImageNode> out;
        const auto& children = track->children();
        const size_t n = children.size();

        // The item whose trimmed range in the track contains the time ...
        size_t index = n;
        OTIO_NS::SerializableObject::Retainer<OTIO_NS::Item> item;
        for (size_t i = 0; i < n; ++i)
        {
            item = OTIO_NS::dynamic_retainer_cast<OTIO_NS::Item>(children[i]);
            if (!item)
            {
                continue;
            }
            const auto range = item->trimmed_range_in_parent();
            if (range.has_value() && range.value().contains(time))
            {
                out = buildItem(track->transformed_time(time, item), item);
                index = i;
                break;
            }
        }
        if (index == n || !item || !out)
        {
            return out;
        }

        // ... blended with its neighbours through the transitions either side of it, when
        // the transition's range contains the time and an item lies beyond it.  The earlier
        // item is always the "from" input.  ImageGraph.cpp:262-328.
        for (int side = -1; side <= 1; side += 2)
        {
            const bool hasNeighbour = side < 0 ? index > 0 : index + 1 < n;
            const bool hasBeyond = side < 0 ? index > 1 : (n > 1 && index + 2 < n);
            if (!hasNeighbour)
            {
                continue;
            }
            auto transition = OTIO_NS::dynamic_retainer_cast<OTIO_NS::Transition>(children[index + side]);
            if (!transition)
            {
                continue;
            }
            const auto range = transition->trimmed_range_in_parent();
            if (!range.has_value() || !range.value().contains(time) || !hasBeyond)
            {
                continue;
            }
            auto other = OTIO_NS::dynamic_retainer_cast<OTIO_NS::Item>(children[index + 2 * side]);
            if (!other)
            {
                continue;
            }
            auto otherNode = buildItem(track->transformed_time(time, other), other);
            out = side < 0 ?
                buildTransition(transition, range.value(), time, otherNode, out) :
                buildTransition(transition, range.value(), time, out, otherNode);
        }
        return out;
    }

    std::shared_ptr<IImageNode> ImageGraph::Frame::buildItem(
        const OTIO_NS::RationalTime& time,
        const OTIO_NS::SerializableObject::Retainer<OTIO_NS::Item>& item)
    {
        std::shared_ptr<IImageNode> out;

        const EffectList& effects = item->effects();
        OTIO_NS::RationalTime t = warpTime(time, item->available_range(), effects);

        if (auto clip = OTIO_NS::dynamic_retainer_cast<OTIO_NS::Clip>(item))
        {
            OTIO_NS::MediaReference* ref = clip->media_reference();
            if (auto generatorRef = dynamic_cast<OTIO_NS::GeneratorReference*>(ref))
            {
                out = host->createNode(generatorRef->parameters(), generatorRef->generator_kind());
            }
            else if (auto read = d.read(ref, true))
            {
                if (dynamic_cast<OTIO_NS::ExternalReference*>(ref) &&
                    clip->available_range().start_time() != read->getTimeRange().start_time())
                {
                    // media without timecode: clip time counts from the available range's start
                    t -= clip->available_range().start_time();
                }
                read->setTime(t);
                out = read;
            }
        }
        else if (OTIO_NS::dynamic_retainer_cast<OTIO_NS::Gap>(item))
        {
            // a gap is a transparent frame
            OTIO_NS::AnyDictionary metaData;
            metaData["size"] = vecToAny(d.imageSize);
            out = host->createNode(metaData, "toucan:Fill");
        }

        out = applyEffects(t, effects, out);

        if (item.value == soloItem)
        {
            soloNode = out;
        }
        return out;
    }

    OTIO_NS::RationalTime ImageGraph::Frame::warpTime(
        const OTIO_NS::RationalTime& time,
        const OTIO_NS::TimeRange& timeRange,
        const EffectList& effects)
    {
        OTIO_NS::RationalTime out = time;
        for (const auto& effect : effects)
        {
            if (auto linearTimeWarp = dynamic_cast<OTIO_NS::LinearTimeWarp*>(effect.value))
            {
                out = OTIO_NS::RationalTime(
                    (out - timeRange.start_time()).value() * linearTimeWarp->time_scalar(),
                    time.rate()).round();
            }
        }
        return out;
    }

    std::shared_ptr<IImageNode> ImageGraph::Frame::applyEffects(
        const OTIO_NS::RationalTime&,
        const EffectList& effects,
        const std::shared_ptr<IImageNode>& input)
    {
        std::shared_ptr<IImageNode> out = input;
        for (const auto& effect : effects)
        {
            // effects the host does not know (including "" of a LinearTimeWarp) are skipped
            if (auto node = host->createNode(effect->metadata(), effect->effect_name(), { out }))
            {
                out = node;
            }
        }
        return out;
    }
}
