// Minimal stand-in for feather-tk's Context / LogSystem / LRUCache (ftk is not available
// in this build environment).  Only what toucanRender uses on the render path:
// lib/toucanRender/ImageEffectHost.cpp:104-218, ImageGraph.cpp:51,63-69.
#pragma once

#include <cstdio>
#include <cstdlib>
#include <list>
#include <map>
#include <memory>
#include <mutex>
#include <string>

namespace ftk {

enum class LogType { Message, Warning, Error };

class LogSystem {
public:
    void print(const std::string& prefix, const std::string& msg, LogType type = LogType::Message)
    {
        static const bool verbose = std::getenv("TOUCAN_LOG") != nullptr;
        if (type == LogType::Message && !verbose) return;
        std::lock_guard<std::mutex> lock(_mutex);
        std::fprintf(stderr, "%s%s: %s\n", type == LogType::Error ? "ERROR: " : (type == LogType::Warning ? "WARNING: " : ""),
                     prefix.c_str(), msg.c_str());
    }

private:
    std::mutex _mutex;
};

class Context : public std::enable_shared_from_this<Context> {
public:
    static std::shared_ptr<Context> create() { return std::shared_ptr<Context>(new Context); }
    template <class T>
    std::shared_ptr<T> getSystem();

private:
    Context() : _log(std::make_shared<LogSystem>()) {}
    std::shared_ptr<LogSystem> _log;
};
template <>
inline std::shared_ptr<LogSystem> Context::getSystem<LogSystem>()
{
    return _log;
}

template <class K, class V>
class LRUCache {
public:
    void setMax(size_t n) { _max = n; }
    bool get(const K& key, V& out)
    {
        auto i = _map.find(key);
        if (i == _map.end()) return false;
        _order.splice(_order.begin(), _order, i->second.second);
        out = i->second.first;
        return true;
    }
    void add(const K& key, const V& value)
    {
        auto i = _map.find(key);
        if (i != _map.end()) {
            i->second.first = value;
            _order.splice(_order.begin(), _order, i->second.second);
            return;
        }
        _order.push_front(key);
        _map[key] = { value, _order.begin() };
        while (_map.size() > _max && !_order.empty()) {
            _map.erase(_order.back());
            _order.pop_back();
        }
    }
    void clear()
    {
        _map.clear();
        _order.clear();
    }

private:
    size_t _max = 10;
    std::list<K> _order;
    std::map<K, std::pair<V, typename std::list<K>::iterator>> _map;
};

} // namespace ftk
