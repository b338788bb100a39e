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

    ImageGraph::ImageGraph(
        const std::shared_ptr<ftk::Context>& context,
        const std::filesystem::path& path,
        const std::shared_ptr<TimelineWrapper>& timelineWrapper) :
        _context(context),
        _path(path),
        _timelineWrapper(timelineWrapper),
        _timeRange(timelineWrapper->getTimeRange())
    {
        _readCache.setMax(20);

        // The timeline's image size / channels / type come from the first video clip that
        // yields them: readable media, or a generator with a "size" parameter (generators
        // are hard-wired to 4 x UINT8).  ImageGraph.cpp:53-123.
        for (const auto& clip : getVideoClips(_timelineWrapper->getTimeline()))
        {
            OTIO_NS::MediaReference* ref = clip->media_reference();
            if (auto generatorRef = dynamic_cast<OTIO_NS::GeneratorReference*>(ref))
            {
                const auto& parameters = generatorRef->parameters();
                const auto i = parameters.find("size");
                if (i != parameters.end() && i->second.has_value())
                {
                    anyToVec(std::any_cast<OTIO_NS::AnyVector>(i->second), _imageSize);
                    _imageChannels = 4;
                    _imageDataType = toImageDataType(TC_U8);
                    break;
                }
            }
            else if (auto read = _read(ref, false))
            {
                if (read->getSpec().width > 0)
                {
                    _imageInfo(read);
                    break;
                }
            }
        }
    }

    ImageGraph::~ImageGraph()
    {}

    const IMATH_NAMESPACE::V2i& ImageGraph::getImageSize() const { return _imageSize; }

    int ImageGraph::getImageChannels() const { return _imageChannels; }

    const std::string& ImageGraph::getImageDataType() const { return _imageDataType; }

    void ImageGraph::_imageInfo(const std::shared_ptr<IReadNode>& read)
    {
        const ImageSpec& spec = read->getSpec();
        _imageSize.x = spec.width;
        _imageSize.y = spec.height;
        _imageChannels = spec.nchannels;
        _imageDataType = toImageDataType(spec.format);
    }

    std::shared_ptr<IReadNode> ImageGraph::_read(const OTIO_NS::MediaReference* ref, bool cache)
    {
        // External and image-sequence references become read nodes; a reader that cannot be
        // opened is logged and yields no node (the clip then contributes nothing).
        std::shared_ptr<IReadNode> read;
        if (!dynamic_cast<const OTIO_NS::ExternalReference*>(ref) &&
            !dynamic_cast<const OTIO_NS::ImageSequenceReference*>(ref))
        {
            return read;
        }
        if (cache && _readCache.get(ref, read))
        {
            return read;
        }
        try
        {
            read = _timelineWrapper->createReadNode(ref);
            if (cache)
            {
                _readCache.add(ref, read);
            }
        }
        catch (const std::exception& e)
        {
            if (auto context = _context.lock())
            {
                context->getSystem<ftk::LogSystem>()->print(logPrefix, e.what(), ftk::LogType::Error);
            }
        }
        return read;
    }

    std::shared_ptr<IImageNode> ImageGraph::exec(
        const std::shared_ptr<ImageEffectHost>& host,
        const OTIO_NS::RationalTime& time,
        const OTIO_NS::Item* itemNode)
    {
        _host = host;
        _itemNode = itemNode;

        // Background: opaque black at the timeline's size.
        OTIO_NS::AnyDictionary metaData;
        metaData["size"] = vecToAny(_imageSize);
        metaData["color"] = vecToAny(IMATH_NAMESPACE::V4f(0.F, 0.F, 0.F, 1.F));
        std::shared_ptr<IImageNode> node = host->createNode(metaData, "toucan:Fill");

        auto stack = _timelineWrapper->getTimeline()->tracks();
        const EffectList& stackEffects = stack->effects();
        OTIO_NS::RationalTime t = time - _timeRange.start_time();
        t = _timeWarps(t, stack->available_range(), stackEffects);

        // Bottom track first; each video track with clips is composited over what is below.
        for (const auto& child : stack->children())
        {
            auto track = OTIO_NS::dynamic_retainer_cast<OTIO_NS::Track>(child);
            if (!track || track->kind() != OTIO_NS::Track::Kind::video || track->find_clips().empty())
            {
                continue;
            }
            const EffectList& trackEffects = track->effects();
            OTIO_NS::RationalTime t2 = t;
            if (!trackEffects.empty())
            {
                t2 = _timeWarps(t2, track->available_range(), trackEffects);
            }
            auto trackNode = _effects(t2, trackEffects, _track(t2, track));

            std::vector<std::shared_ptr<IImageNode> > nodes;
            if (trackNode) nodes.push_back(trackNode);
            if (node) nodes.push_back(node);
            auto comp = std::make_shared<CompNode>(nodes);
            comp->setPremult(true);
            node = comp;
        }

        node = _effects(t, stackEffects, node);

        _host.reset();
        _itemNode = nullptr;
        if (_outNode)
        {
            node = _outNode;
        }
        _outNode.reset();
        return node;
    }

    std::shared_ptr<IImageNode> ImageGraph::_transition(
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
        auto node = _host->createNode(metaData, transition->transition_type(), { from, to });
        if (!node)
        {
            // unknown transition types fall back to a dissolve
            node = _host->createNode(metaData, "toucan:Dissolve", { from, to });
        }
        return node;
    }

    std::shared_ptr<IImageNode> ImageGraph::_track(
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
                out = _item(track->transformed_time(time, item), item);
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
            auto otherNode = _item(track->transformed_time(time, other), other);
            out = side < 0 ?
                _transition(transition, range.value(), time, otherNode, out) :
                _transition(transition, range.value(), time, out, otherNode);
        }
        return out;
    }

    std::shared_ptr<IImageNode> ImageGraph::_item(
        const OTIO_NS::RationalTime& time,
        const OTIO_NS::SerializableObject::Retainer<OTIO_NS::Item>& item)
    {
        std::shared_ptr<IImageNode> out;

        const EffectList& effects = item->effects();
        OTIO_NS::RationalTime t = _timeWarps(time, item->available_range(), effects);

        if (auto clip = OTIO_NS::dynamic_retainer_cast<OTIO_NS::Clip>(item))
        {
            OTIO_NS::MediaReference* ref = clip->media_reference();
            if (auto generatorRef = dynamic_cast<OTIO_NS::GeneratorReference*>(ref))
            {
                out = _host->createNode(generatorRef->parameters(), generatorRef->generator_kind());
            }
            else if (auto read = _read(ref, true))
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
            metaData["size"] = vecToAny(_imageSize);
            out = _host->createNode(metaData, "toucan:Fill");
        }

        out = _effects(t, effects, out);

        if (item.value == _itemNode)
        {
            _outNode = out;
        }
        return out;
    }

    OTIO_NS::RationalTime ImageGraph::_timeWarps(
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

    std::shared_ptr<IImageNode> ImageGraph::_effects(
        const OTIO_NS::RationalTime&,
        const EffectList& effects,
        const std::shared_ptr<IImageNode>& input)
    {
        std::shared_ptr<IImageNode> out = input;
        for (const auto& effect : effects)
        {
            // effects the host does not know (including "" of a LinearTimeWarp) are skipped
            if (auto node = _host->createNode(effect->metadata(), effect->effect_name(), { out }))
            {
                out = node;
            }
        }
        return out;
    }
}
