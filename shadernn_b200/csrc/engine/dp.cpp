// dp::loadFromJsonModel + dp::generateInferenceGraph (counterparts of core/src/ic2/dp.cpp:115-167 and :389-640).
#include <deque>
#include <unordered_map>

#include "engine.h"

namespace snn {
namespace dp {

static std::string baseName(const std::string& path) {
    size_t pos = path.find_last_of('/');
    return pos == std::string::npos ? path : path.substr(pos + 1);
}

std::vector<std::shared_ptr<GenericModelLayer>> loadFromJsonModel(const std::string& fileName) {
    std::vector<std::shared_ptr<GenericModelLayer>> layers;
    ModelParser parser(fileName);
    const int layerCount = parser.getLayerCount();
    initLayerRegisty();
    const std::string shortName = baseName(fileName);
    for (int i = 0; i < layerCount; i++) {
        const std::string layerName = parser.getLayerName(i);
        GenericModelLayer* L        = createLayerInstance(layerName, parser, i);
        layers.emplace_back(std::shared_ptr<GenericModelLayer>(L));
        char buf[32];
        snprintf(buf, sizeof(buf), " layer [%02d] ", i);
        L->name = shortName + buf + layerName; // dp.cpp:135: "%s layer [%02d] %s" with the JSON's own type string
    }
    if (layers.empty()) throw std::runtime_error("head layer not found.");
    // build layer connections (dp.cpp:145-155)
    for (int i = 0; i < layerCount; i++) {
        for (int ii : parser.getInboundLayerId(i)) {
            if (ii < 0 || ii >= layerCount) throw std::runtime_error("layer " + std::to_string(i) + ": inputId " + std::to_string(ii) + " out of range");
            layers[i]->prevLayers.push_back(layers[ii].get());
            layers[ii]->nextLayers.push_back(layers[i].get());
        }
    }
    return layers;
}

// Kahn's algorithm (dp.cpp:389-429). Ready nodes are taken in JSON order so the execution order is stable.
static std::vector<GenericModelLayer*> topologicalSort2(const std::vector<std::shared_ptr<GenericModelLayer>>& layers) {
    std::vector<GenericModelLayer*> sorted;
    std::unordered_map<GenericModelLayer*, size_t> ready;
    std::deque<GenericModelLayer*> pending;
    for (auto& l : layers)
        if (l->prevLayers.empty()) pending.push_back(l.get());
    while (!pending.empty()) {
        GenericModelLayer* node = pending.front();
        pending.pop_front();
        sorted.push_back(node);
        for (auto* next : node->nextLayers)
            if (++ready[next] == next->prevLayers.size()) pending.push_back(next);
    }
    if (sorted.size() != layers.size()) throw std::runtime_error("Not a DAG - cycle in graph or incorrect number of input nodes"); // dp.cpp:90
    return sorted;
}

InferenceGraph generateInferenceGraph(const std::vector<std::shared_ptr<GenericModelLayer>>& layers, const ShaderGenOptions& options) {
    InferenceGraph g;
    g.sorted = topologicalSort2(layers);
    std::unordered_map<GenericModelLayer*, Dims> out;
    for (auto* L : g.sorted) {
        L->inputDims.clear();
        if (L->isInputLayer) {
            auto* in = static_cast<InputLayerLayer*>(L);
            Dims d;
            d.width  = options.desiredInputWidth ? options.desiredInputWidth : in->_desc.inputWidth;
            d.height = options.desiredInputHeight ? options.desiredInputHeight : in->_desc.inputHeight;
            d.depth  = in->_desc.inputChannels;
            if (!d.width || !d.height) throw std::runtime_error(L->name + ": input dimensions unknown (set input_width/height)");
            L->inputDims.push_back(d);
        } else {
            for (auto* p : L->prevLayers) L->inputDims.push_back(out.at(p));
            if (L->inputDims.empty()) throw std::runtime_error(L->name + ": layer has no inputs");
        }
        Dims o;
        L->getOutputDims(o.width, o.height, o.depth);
        if (!o.width || !o.height || !o.depth) throw std::runtime_error(L->name + ": empty output dimensions");
        out[L] = o;
        g.outputDims.push_back(o);
    }
    return g;
}

} // namespace dp
} // namespace snn
