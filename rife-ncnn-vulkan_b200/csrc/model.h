// model.h -- loader for the reference's model-directory layout (ncnn .param text graph + .bin weights).
// Format facts follow /root/reference/src/ncnn/src/net.cpp:1374-1590 (param), paramdict.cpp:151-330
// (key=value / array syntax, float parsing) and modelbin.cpp:89-260 (weight tags); see SURVEY.md 3.5.
#pragma once
#include <map>
#include <string>
#include <vector>

namespace rife {

struct ParamVal {
    bool is_array = false;
    bool is_float = false;
    int i = 0;
    float f = 0.f;
    std::vector<float> af;  // arrays keep both views
    std::vector<int> ai;
};

struct Layer {
    std::string type, name;
    std::vector<int> bottoms, tops;  // blob ids
    std::map<int, ParamVal> params;
    // weights (host, fp32 values; conv weights are exactly representable in fp16)
    std::vector<float> weight, bias, slope;
    bool weight_is_fp16 = false;

    int geti(int id, int def) const {
        auto it = params.find(id);
        if (it == params.end()) return def;
        return it->second.is_float ? (int)it->second.f : it->second.i;
    }
    float getf(int id, float def) const {
        auto it = params.find(id);
        if (it == params.end()) return def;
        return it->second.is_float ? it->second.f : (float)it->second.i;
    }
    const ParamVal* get(int id) const {
        auto it = params.find(id);
        return it == params.end() ? nullptr : &it->second;
    }
};

struct Net {
    std::vector<Layer> layers;
    std::vector<std::string> blob_names;
    std::vector<int> producer;  // blob id -> layer index
    int find_blob(const std::string& n) const {
        for (size_t i = 0; i < blob_names.size(); i++)
            if (blob_names[i] == n) return (int)i;
        return -1;
    }
};

// returns 0 on success; negative on error (message in err)
int load_net(const std::string& param_path, const std::string& bin_path, Net& net, std::string& err);
// same, from memory (tag = name used in error messages)
int parse_net(const std::string& param_text, const std::string& bin_bytes, const std::string& tag, Net& net, std::string& err);

}  // namespace rife
