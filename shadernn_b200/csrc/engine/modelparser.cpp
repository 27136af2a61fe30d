// ModelParser + per-op desc parsing for the SNN JSON model format (reader counterpart of
// core/src/ic2/modelparser.cpp; the format is reconstructed in SURVEY.md Appendix B).
#include <algorithm>
#include <cstdio>
#include <fstream>
#include <sstream>

#include "engine.h"

namespace snn {
namespace dp {

static std::string dirOf(const std::string& path) {
    size_t pos = path.find_last_of('/');
    return pos == std::string::npos ? std::string(".") : path.substr(0, pos);
}

ModelParser::ModelParser(const std::string& fileName): _fileName(fileName) {
    std::ifstream f(fileName, std::ios::binary);
    if (!f) throw std::runtime_error("ModelParser:: Could not load JSON file " + fileName); // modelparser.cpp:224-226
    std::stringstream ss;
    ss << f.rdbuf();
    std::string text = ss.str();
    try {
        _root = json::parse(text);
    } catch (std::exception& e) { throw std::runtime_error("ModelParser:: Could not parse JSON file " + fileName + ": " + e.what()); }
    if (!_root->isObject() || !_root->has("numLayers")) throw std::runtime_error("ModelParser:: no numLayers in " + fileName);
    const json::Value& numNode = _root->at("numLayers");
    if (numNode.has("bin_file_name")) { // modelparser.cpp:235-257 (the reference resolves it under MODEL_DIR; we look next to the JSON)
        std::string bin = dirOf(fileName) + "/" + numNode.at("bin_file_name").asString();
        binFile         = fopen(bin.c_str(), "rb");
        if (!binFile) throw std::runtime_error("ModelParser:: open " + bin + " failed");
    }
}

ModelParser::~ModelParser() {
    if (binFile) fclose(binFile);
}

int ModelParser::getLayerCount() const { return (int) _root->at("numLayers").at("count").asNumber(); }

const json::Value& ModelParser::layer(int layerId) const { return _root->at("Layer_" + std::to_string(layerId)); }

std::string ModelParser::getLayerName(int layerId) const {
    const json::Value& l = layer(layerId);
    std::string cls      = l.at("type").asString();
    if (cls == "Lambda") cls = l.at("name").asString();
    return cls;
}

int ModelParser::getNumInbound(int layerId) const { return (int) layer(layerId).at("numInputs").asNumber(); }

std::vector<int> ModelParser::getInboundLayerId(int layerId) const {
    const int numIn = getNumInbound(layerId);
    std::vector<int> ids;
    if (numIn == 0) return ids;
    const json::Value& nodes = layer(layerId).at("inputId");
    for (int i = 0; i < numIn; ++i) ids.push_back((int) nodes.numAt(i));
    return ids;
}

int ModelParser::getInputPlanes(int layerId) const {
    if (getNumInbound(layerId) == 0) return 0;
    return (int) layer(layerId).at("inputPlanes").asNumber();
}
int ModelParser::getOutputPlanes(int layerId) const { return (int) layer(layerId).at("outputPlanes").asNumber(); }

bool ModelParser::isInputRange01() const {
    const json::Value* r = _root->find("inputRange");
    return r && r->isString() && r->str == "[0,1]";
}

void ModelParser::readBin(float* dst, size_t count) {
    if (!binFile) throw std::runtime_error("ModelParser:: no sidecar .bin is open");
    if (fread(dst, sizeof(float), count, binFile) != count) throw std::runtime_error("ModelParser:: sidecar .bin is truncated");
}

// ---- padding ----------------------------------------------------------------------------------------------
// modelparser.cpp:584-609 (conv), :338-363 (maxpool): "padding" is [[t,b],[l,r]] (+ "mode" for conv), [t,l], a number
// or a string.
void PaddingSpec::parse(const json::Value& layerObj, bool readMode) {
    const json::Value* pv = layerObj.find("padding");
    if (!pv) throw std::runtime_error("missing key 'padding'");
    auto u = [](double d) { return std::to_string((uint32_t) d); };
    if (pv->isArray()) {
        if (pv->type == json::Value::Array && pv->size() >= 2 && pv->elemAt(0).isArray()) {
            t = u(pv->elemAt(0).numAt(0));
            b = u(pv->elemAt(0).numAt(1));
            l = u(pv->elemAt(1).numAt(0));
            r = u(pv->elemAt(1).numAt(1));
            if (readMode) mode = layerObj.at("mode").asString(); // conv only; required in this form (modelparser.cpp:594)
        } else {
            t = u(pv->numAt(0));
            l = u(pv->numAt(1));
            b = t;
            r = l;
        }
    } else if (pv->isNumber()) {
        t = b = l = r = u(pv->num);
    } else {
        t = b = l = r = pv->asString();
    }
}

void PaddingSpec::offsets(int kernelSize, bool evenMinusOne, uint32_t (&offs)[4]) const {
    const bool isdigit = !t.empty() && std::all_of(t.begin(), t.end(), ::isdigit);
    if (isdigit) {
        offs[0] = (uint32_t) std::stoul(t);
        offs[1] = (uint32_t) std::stoul(b);
        offs[2] = (uint32_t) std::stoul(l);
        offs[3] = (uint32_t) std::stoul(r);
        return;
    }
    offs[0] = offs[1] = offs[2] = offs[3] = 0;
    if (t == "valid" || t == "none") return;
    if (kernelSize > 1) {
        const uint32_t p = std::max((uint32_t) (kernelSize / 2), (uint32_t) 1);
        offs[0] = offs[1] = offs[2] = offs[3] = p;
        if (evenMinusOne && kernelSize % 2 == 0) {
            offs[0] -= 1;
            offs[2] -= 1;
        }
    }
}

int ActivationSpec::fromString(const std::string& s) {
    if (s == "relu") return SNNB_ACT_RELU;
    if (s == "relu6") return SNNB_ACT_RELU6;
    if (s == "tanh") return SNNB_ACT_TANH;
    if (s == "sigmoid") return SNNB_ACT_SIGMOID;
    if (s == "leakyRelu" || s == "leaky_relu") return SNNB_ACT_LEAKY_RELU;
    if (s == "SiLU" || s == "silu" || s == "swish") return SNNB_ACT_SILU;
    if (s == "softmax") return SNNB_ACT_SOFTMAX;
    return SNNB_ACT_NONE; // "linear", "identity", "", anything else
}

static float leakyAlpha(const json::Value& l, float dflt, bool required) {
    if (l.has("leakyReluAlpha")) return (float) l.at("leakyReluAlpha").asNumber();
    if (l.has("alpha")) return (float) l.at("alpha").asNumber();
    if (required) throw std::runtime_error("leakyRelu without leakyReluAlpha/alpha");
    return dflt;
}

static void readFloats(ModelParser& parser, const json::Value* arr, size_t count, std::vector<float>& dst, const char* what, bool allowBin = true) {
    dst.resize(count);
    if (parser.isBinWeight() && allowBin) {
        parser.readBin(dst.data(), count);
        return;
    }
    if (!arr || !arr->isArray() || arr->size() < count) throw std::runtime_error(std::string("weights array '") + what + "' missing or too short");
    for (size_t i = 0; i < count; ++i) dst[i] = (float) arr->numAt(i);
}

static bool isTrue(const json::Value& l, const char* key) {
    const json::Value* v = l.find(key);
    return v && v->isString() && v->str == "True";
}

static void parseBN(ModelParser& parser, const json::Value& l, int C, std::map<std::string, std::vector<float>>& bn, bool defaultsAllowed, bool allowBin = true) {
    // order in the sidecar stream: gamma, beta, mean, var (modelparser.cpp:694-726)
    const json::Value* o = l.find("batchNormalization");
    std::vector<float> g, b, m, v;
    if (parser.isBinWeight() && allowBin) {
        readFloats(parser, nullptr, C, g, "gamma");
        readFloats(parser, nullptr, C, b, "beta");
        readFloats(parser, nullptr, C, m, "moving_mean");
        readFloats(parser, nullptr, C, v, "moving_variance");
    } else {
        if (!o) throw std::runtime_error("missing key 'batchNormalization'");
        const json::Value* jb = o->find("beta");
        const json::Value* jg = o->find("gamma");
        if (jb)
            readFloats(parser, jb, C, b, "beta", false);
        else if (defaultsAllowed)
            b.assign(C, 0.0f); // modelparser.cpp:1047-1078
        else
            throw std::runtime_error("missing key 'beta'");
        if (jg)
            readFloats(parser, jg, C, g, "gamma", false);
        else if (defaultsAllowed)
            g.assign(C, 1.0f);
        else
            throw std::runtime_error("missing key 'gamma'");
        const json::Value* jm = o->find("moving_mean");
        if (!jm) jm = o->find("movingMean");
        const json::Value* jv = o->find("moving_variance");
        if (!jv) jv = o->find("movingVariance");
        readFloats(parser, jm, C, m, "moving_mean", false);
        readFloats(parser, jv, C, v, "moving_variance", false);
    }
    bn["gamma"] = g, bn["beta"] = b, bn["movingMean"] = m, bn["movingVariance"] = v;
}

// ---- InputLayer (modelparser.cpp:480-497) ----
void InputLayerDesc::parse(ModelParser& parser, int layerId) {
    const json::Value& l = parser.layer(layerId);
    const json::Value* w = l.find("Input Width");
    if (!w) w = l.find("Input Weight"); // the ONNX converter's typo (tools/convertTool/.../input.py:57)
    const json::Value* h = l.find("Input Height");
    inputWidth           = w ? (uint32_t) w->asNumber() : 0;
    inputHeight          = h ? (uint32_t) h->asNumber() : 0;
    inputChannels        = (uint32_t) l.at("outputPlanes").asNumber();
    if (l.has("inputIndex")) inputIndex = (uint32_t) l.at("inputIndex").asNumber();
}

// ---- Conv2D (modelparser.cpp:574-781) ----
void Conv2DDesc::parse(ModelParser& parser, int layerId) {
    const json::Value& l = parser.layer(layerId);
    const int OC = (int) l.at("outputPlanes").asNumber(), IC = (int) l.at("inputPlanes").asNumber();
    activationName = l.at("activation").asString();
    activation.id  = ActivationSpec::fromString(activationName);
    padding.parse(l, true);
    kernelSize = (uint32_t) l.at("kernel_size").asNumber();
    stride     = (uint32_t) l.at("strides").asNumber();
    const json::Value* wobj = l.find("weights");
    const size_t wcount     = (size_t) OC * IC * kernelSize * kernelSize;
    readFloats(parser, wobj ? wobj->find("kernel") : nullptr, wcount, weights, "kernel"); // OIHW
    if (isTrue(l, "useBias")) readFloats(parser, wobj ? wobj->find("bias") : nullptr, OC, biases, "bias");
    useBatchNormalization = isTrue(l, "useBatchNormalization");
    if (useBatchNormalization) parseBN(parser, l, OC, batchNormalization, false);
    if (activationName == "leakyRelu") activation.alpha = leakyAlpha(l, 0.0f, true);
}

// ---- Depthwise (modelparser.cpp:783-985): JSON kernel is [kh*kw][C] (HWC-major), the .bin variant is [C][kh][kw] ----
void SeparableConv2DDesc::parse(ModelParser& parser, int layerId) {
    const json::Value& l = parser.layer(layerId);
    const int C = (int) l.at("inputPlanes").asNumber();
    activationName = l.has("activation") ? l.at("activation").asString() : "";
    activation.id  = ActivationSpec::fromString(activationName);
    padding.parse(l, false);
    kernelSize = (uint32_t) l.at("kernel_size").asNumber();
    stride     = (uint32_t) l.at("strides").asNumber();
    const json::Value* wobj = l.find("weights");
    const size_t plane      = (size_t) kernelSize * kernelSize;
    if (parser.isBinWeight()) {
        readFloats(parser, nullptr, plane * C, weights, "kernel");
    } else {
        std::vector<float> hwc;
        readFloats(parser, wobj ? wobj->find("kernel") : nullptr, plane * C, hwc, "kernel");
        weights.resize(plane * C);
        for (size_t i = 0; i < plane; ++i)
            for (int c = 0; c < C; ++c) weights[(size_t) c * plane + i] = hwc[i * C + c]; // modelparser.cpp:843-850
    }
    if (isTrue(l, "useBias")) readFloats(parser, wobj ? wobj->find("bias") : nullptr, C, biases, "bias");
    useBatchNormalization = isTrue(l, "useBatchNormalization");
    if (useBatchNormalization) parseBN(parser, l, C, batchNormalization, false);
    if (activationName == "leakyRelu") activation.alpha = leakyAlpha(l, 0.0f, true);
}

// ---- pools ----
static uint32_t numberOrFirst(const json::Value& v) { return (uint32_t) (v.isArray() ? v.numAt(0) : v.asNumber()); }

void PoolDesc::parseMax(ModelParser& parser, int layerId) { // modelparser.cpp:304-371
    const json::Value& l = parser.layer(layerId);
    const json::Value* pool = l.find("pool");
    if (!pool) pool = l.find("pool_size");
    if (!pool) throw std::runtime_error("missing key 'pool'");
    kernelSize = numberOrFirst(*pool);
    if (l.has("stride"))
        stride = numberOrFirst(l.at("stride"));
    else if (l.has("strides"))
        stride = numberOrFirst(l.at("strides"));
    else
        stride = kernelSize;
    padding.parse(l, false);
}
void PoolDesc::parseAvg(ModelParser& parser, int layerId) { // modelparser.cpp:373-397
    const json::Value& l = parser.layer(layerId);
    const json::Value* pool = l.find("pool");
    if (!pool) pool = l.find("pool_size");
    if (!pool) throw std::runtime_error("missing key 'pool'/'pool_size'");
    kernelSize = numberOrFirst(*pool);
    // Only "stride" is read here; the converter's "strides" key is ignored and the stride then defaults to the pool
    // size (modelparser.cpp:385-391) - which is what turns the converter's global pool (pool_size=[H,H], strides=[1,1],
    // averagepooling2d.py:40-55) into a 1x1 output.
    if (l.has("stride"))
        stride = numberOrFirst(l.at("stride"));
    else
        stride = kernelSize;
    const json::Value& p = l.at("padding");
    padding.t = padding.b = padding.l = padding.r = p.isString() ? p.str : std::to_string((uint32_t) p.asNumber());
}

// ---- layer creators ------------------------------------------------------------------------------------------
static void common(GenericModelLayer* L, ModelParser& parser, int i) {
    L->numOutputPlanes = (uint32_t) parser.getOutputPlanes(i);
    L->numInputPlanes  = (uint32_t) parser.getInputPlanes(i);
}

static GenericModelLayer* InputLayerCreator(ModelParser& parser, int i) {
    auto* L = new InputLayerLayer();
    L->_desc.parse(parser, i);
    L->isInputLayer    = true;
    L->numOutputPlanes = L->_desc.inputChannels;
    return L;
}
static GenericModelLayer* Conv2DCreator(ModelParser& parser, int i) {
    auto* L = new Conv2DLayer();
    common(L, parser, i);
    L->_desc.parse(parser, i);
    return L;
}
static GenericModelLayer* SeparableConv2DCreator(ModelParser& parser, int i) {
    auto* L = new SeparableConv2DLayer();
    common(L, parser, i);
    L->_desc.parse(parser, i);
    return L;
}
static GenericModelLayer* MaxPooling2DCreator(ModelParser& parser, int i) {
    auto* L = new PoolingLayer();
    common(L, parser, i);
    L->_desc.parseMax(parser, i);
    return L;
}
static GenericModelLayer* AveragePooling2DCreator(ModelParser& parser, int i) {
    auto* L = new PoolingLayer();
    common(L, parser, i);
    L->isAvg = true;
    L->_desc.parseAvg(parser, i);
    return L;
}
static GenericModelLayer* AdaptiveAvgPool2dCreator(ModelParser& parser, int i) {
    auto* L = new AdaptiveAvgPool2dLayer();
    common(L, parser, i);
    L->poolSize = numberOrFirst(parser.layer(i).at("pool")); // modelparser.cpp:439-450
    return L;
}
static void parseActivation(const json::Value& l, ActivationSpec& a, const char* leakyName, float dfltAlpha) {
    std::string s = l.has("activation") ? l.at("activation").asString() : "linear";
    a.id          = ActivationSpec::fromString(s);
    if (s == leakyName || a.id == SNNB_ACT_LEAKY_RELU) a.alpha = leakyAlpha(l, dfltAlpha, false);
}
static GenericModelLayer* AddCreator(ModelParser& parser, int i) { // modelparser.cpp:399-419
    auto* L = new AddLayer();
    common(L, parser, i);
    parseActivation(parser.layer(i), L->activation, "leaky_relu", 0.3f);
    return L;
}
static GenericModelLayer* ActivationCreator(ModelParser& parser, int i) { // modelparser.cpp:421-437
    auto* L = new ActivationLayer();
    common(L, parser, i);
    parseActivation(parser.layer(i), L->activation, "leaky_relu", 0.3f);
    return L;
}
static GenericModelLayer* BatchNormalizationCreator(ModelParser& parser, int i) { // modelparser.cpp:1011-1110
    auto* L = new BatchNormalizationLayer();
    common(L, parser, i);
    parseBN(parser, parser.layer(i), (int) L->numOutputPlanes, L->batchNormalization, true, /*allowBin=*/false); // always JSON (modelparser.cpp:1011-1110 has no .bin branch)
    parseActivation(parser.layer(i), L->activation, "leakyRelu", 0.0f);
    return L;
}
static GenericModelLayer* InstanceNormCreator(ModelParser& parser, int i) { // modelparser.cpp:1149-1201
    auto* L = new InstanceNormLayer();
    common(L, parser, i);
    const json::Value& l = parser.layer(i);
    const json::Value& w = l.at("weights");
    const int C          = (int) L->numOutputPlanes;
    // always embedded (the reference reads weights{bias,scale} from JSON even for .bin models)
    L->beta.resize(C), L->gamma.resize(C);
    for (int c = 0; c < C; ++c) L->beta[c] = (float) w.at("bias").numAt(c), L->gamma[c] = (float) w.at("scale").numAt(c);
    parseActivation(l, L->activation, "leakyRelu", 0.0f);
    return L;
}
static GenericModelLayer* DenseCreator(ModelParser& parser, int i) { // modelparser.cpp:499-572
    auto* L = new DenseLayer();
    common(L, parser, i);
    const json::Value& l = parser.layer(i);
    L->units             = l.has("units") ? (uint32_t) l.at("units").asNumber() : L->numOutputPlanes;
    if (L->units == 0) throw std::runtime_error("Dense layer " + std::to_string(i) + ": units must be positive");
    const json::Value* wobj = l.find("weights");
    if (parser.isBinWeight()) {
        readFloats(parser, nullptr, (size_t) L->numInputPlanes * L->units, L->kernel, "kernel");
    } else {
        const json::Value* k = wobj ? wobj->find("kernel") : nullptr;
        if (!k || !k->isArray()) throw std::runtime_error("Dense: missing weights.kernel");
        readFloats(parser, k, k->size(), L->kernel, "kernel");
        L->numInputPlanes = (uint32_t) (k->size() / L->units); // numInputUnits = size / units (modelparser.cpp:527)
    }
    if (isTrue(l, "useBias"))
        readFloats(parser, wobj ? wobj->find("bias") : nullptr, L->units, L->biases, "bias");
    std::string s    = l.at("activation").asString();
    L->activation.id = ActivationSpec::fromString(s);
    if (L->activation.id == SNNB_ACT_LEAKY_RELU) L->activation.alpha = leakyAlpha(l, 0.3f, false);
    return L;
}
static GenericModelLayer* FlattenCreator(ModelParser& parser, int i) {
    auto* L = new FlattenLayer();
    common(L, parser, i);
    parseActivation(parser.layer(i), L->activation, "leaky_relu", 0.3f);
    return L;
}
static GenericModelLayer* ConcatenateCreator(ModelParser& parser, int i) {
    auto* L = new ConcatenateLayer();
    common(L, parser, i);
    return L;
}
static GenericModelLayer* UpSampling2DCreator(ModelParser& parser, int i) { // modelparser.cpp:987-1009
    auto* L = new UpSampling2DLayer();
    common(L, parser, i);
    const json::Value& l  = parser.layer(i);
    L->scale              = (float) l.at("scaleFactor").asNumber();
    L->interpolationType  = l.at("interpolation").asString();
    return L;
}
static GenericModelLayer* PadCreator(ModelParser& parser, int i) { // modelparser.cpp:1112-1147
    auto* L = new PadLayer();
    common(L, parser, i);
    const json::Value& l = parser.layer(i);
    auto u               = [](double d) { return std::to_string((uint32_t) d); };
    if (l.has("pads")) { // ONNX order: T = pads[2], L = pads[3], B = pads[6], R = pads[7]
        const json::Value& p = l.at("pads");
        L->padding.t = u(p.numAt(2)), L->padding.b = u(p.numAt(6)), L->padding.l = u(p.numAt(3)), L->padding.r = u(p.numAt(7));
    } else {
        L->padding.parse(l, false);
    }
    // Extension over the reference reader (which drops "mode", leaving every Pad constant): honour an explicit
    // "mode" so that converted style-transfer graphs with ReflectionPad work as the shader supports (vk_pad.comp:53-66).
    if (l.has("mode") && l.at("mode").isString()) L->mode = l.at("mode").str;
    return L;
}
static GenericModelLayer* SubpixelCreator(ModelParser& parser, int i) {
    auto* L = new SubpixelLayer();
    common(L, parser, i);
    const json::Value& l = parser.layer(i);
    if (l.has("kernel_size")) L->kernelSize = (uint32_t) l.at("kernel_size").asNumber();
    return L;
}
static GenericModelLayer* YOLOCreator(ModelParser& parser, int i) {
    auto* L = new YOLOLayer();
    common(L, parser, i);
    return L;
}

static std::map<std::string, LayerCreator>& registry() {
    static std::map<std::string, LayerCreator> r;
    return r;
}
void registerLayer(const std::string& layerName, LayerCreator creator) { registry()[layerName] = creator; }

void initLayerRegisty() { // layerFactory.cpp:109-129 (Conv2DTranspose, Calculate, Unary: out of scope — SURVEY §2)
    if (!registry().empty()) return;
    registerLayer("InputLayer", InputLayerCreator);
    registerLayer("Conv2D", Conv2DCreator);
    registerLayer("Subpixel", SubpixelCreator);
    registerLayer("Concatenate", ConcatenateCreator);
    registerLayer("UpSampling2D", UpSampling2DCreator);
    registerLayer("Add", AddCreator);
    registerLayer("SeparableConv2D", SeparableConv2DCreator);
    registerLayer("Dense", DenseCreator);
    registerLayer("MaxPooling2D", MaxPooling2DCreator);
    registerLayer("AveragePooling2D", AveragePooling2DCreator);
    registerLayer("AdaptiveAvgPool2d", AdaptiveAvgPool2dCreator);
    registerLayer("Flatten", FlattenCreator);
    registerLayer("Pad", PadCreator);
    registerLayer("BatchNormalization", BatchNormalizationCreator);
    registerLayer("InstanceNorm", InstanceNormCreator);
    registerLayer("YOLO", YOLOCreator);
    registerLayer("Activation", ActivationCreator); // creatable-only in the reference (layerFactory.h:125-148)
}

// ---- layers registered through the C-ABI (snnb_register_layer) ----
struct PluginEntry {
    snnb_layer_creator creator;
    void* user;
};
static std::map<std::string, PluginEntry>& pluginRegistry() {
    static std::map<std::string, PluginEntry> r;
    return r;
}
void PluginLayer::getOutputDims(uint32_t& w, uint32_t& h, uint32_t& d) const {
    std::vector<int> in;
    for (auto& dim : inputDims) in.push_back((int) dim.height), in.push_back((int) dim.width), in.push_back((int) dim.depth);
    int out[3] = {0, 0, 0};
    if (!impl.output_dims || impl.output_dims(impl.user, (int) inputDims.size(), in.data(), out) || out[0] <= 0 || out[1] <= 0 || out[2] <= 0)
        throw std::runtime_error(name + ": the registered layer's output_dims callback failed");
    h = (uint32_t) out[0], w = (uint32_t) out[1], d = (uint32_t) out[2];
}
int PluginLayer::run(snnb_context* ctx, const ExecOptions&) {
    std::vector<const snnb_tensor*> ins(inputs.begin(), inputs.end());
    if (!impl.run || impl.run(impl.user, ctx, (int) ins.size(), ins.data(), output)) {
        if (!*snnb::get_error()) snnb::set_error("%s: the registered layer's run callback failed", name.c_str());
        return 1;
    }
    return 0;
}

GenericModelLayer* createLayerInstance(std::string layerName, ModelParser& parser, int i) { // layerFactory.cpp:136-159
    {
        auto pit = pluginRegistry().find(layerName);
        if (pit != pluginRegistry().end()) {
            auto* L = new PluginLayer();
            common(L, parser, i);
            if (pit->second.creator(pit->second.user, reinterpret_cast<const snnb_layer_json*>(&parser.layer(i)), &L->impl) || !L->impl.output_dims || !L->impl.run) {
                delete L;
                throw std::runtime_error("ModelParser: the creator registered for layer type '" + layerName + "' failed on layer " + std::to_string(i));
            }
            L->typeName = layerName;
            L->layerId  = i;
            return L;
        }
    }
    if (layerName == "DepthwiseConv2D" || layerName == "Depthwise") layerName = "SeparableConv2D";
    if (layerName == "InstanceNormalization") layerName = "InstanceNorm";
    if (layerName == "ZeroPadding2D") layerName = "Pad";
    if (layerName == "subpixel" || layerName == "depth_to_space") layerName = "Subpixel";
    auto it = registry().find(layerName);
    if (it == registry().end()) throw std::runtime_error("Not found layer: " + layerName);
    GenericModelLayer* L = nullptr;
    try {
        L = it->second(parser, i);
    } catch (std::exception& e) {
        throw std::runtime_error("ModelParser: issues parsing layer " + std::to_string(i) + " (" + layerName + "): " + e.what());
    }
    L->typeName = layerName;
    L->layerId  = i;
    return L;
}

} // namespace dp
} // namespace snn

// ---- C-ABI: layer registration + read-only JSON accessors (include/snnb.h) ----
extern "C" {
int snnb_register_layer(const char* type_name, snnb_layer_creator creator, void* registry_user) {
    SNNB_REQUIRE(type_name && *type_name && creator, "snnb_register_layer: null argument");
    snn::dp::pluginRegistry()[type_name] = snn::dp::PluginEntry {creator, registry_user};
    return 0;
}
int snnb_unregister_layer(const char* type_name) {
    SNNB_REQUIRE(type_name, "snnb_unregister_layer: null argument");
    SNNB_REQUIRE(snn::dp::pluginRegistry().erase(type_name) == 1, "snnb_unregister_layer: '%s' is not registered", type_name);
    return 0;
}
static const snn::json::Value* jsonAt(const snnb_layer_json* layer, const char* path) {
    const snn::json::Value* v = reinterpret_cast<const snn::json::Value*>(layer);
    std::string p(path ? path : "");
    size_t pos = 0;
    while (v && pos <= p.size()) {
        const size_t dot = p.find('.', pos);
        const std::string key = p.substr(pos, dot == std::string::npos ? std::string::npos : dot - pos);
        v = v->isObject() ? v->find(key) : nullptr;
        if (dot == std::string::npos) break;
        pos = dot + 1;
    }
    return v;
}
int snnb_layer_json_number(const snnb_layer_json* layer, const char* key, double* out) {
    SNNB_REQUIRE(layer && key && out, "snnb_layer_json_number: null argument");
    const snn::json::Value* v = jsonAt(layer, key);
    if (!v || !v->isNumber()) return 1;
    *out = v->num;
    return 0;
}
int snnb_layer_json_string(const snnb_layer_json* layer, const char* key, char* buf, int cap) {
    SNNB_REQUIRE(layer && key && buf && cap > 0, "snnb_layer_json_string: bad argument");
    const snn::json::Value* v = jsonAt(layer, key);
    if (!v || !v->isString()) return 1;
    strncpy(buf, v->str.c_str(), (size_t) cap - 1);
    buf[cap - 1] = 0;
    return 0;
}
int snnb_layer_json_numbers(const snnb_layer_json* layer, const char* path, const double** data, size_t* count) {
    SNNB_REQUIRE(layer && path && data && count, "snnb_layer_json_numbers: null argument");
    const snn::json::Value* v = jsonAt(layer, path);
    if (!v || v->type != snn::json::Value::NumArray) return 1;
    *data = v->nums.data(), *count = v->nums.size();
    return 0;
}
} // extern "C"
