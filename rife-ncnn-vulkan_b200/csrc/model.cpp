// model.cpp -- see model.h.  Written from the format description, not from ncnn's parser code.
#include "model.h"

#include <ctype.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <fstream>
#include <sstream>

namespace rife {

static bool token_is_float(const std::string& s) {
    for (char c : s)
        if (c == '.' || c == 'e' || c == 'E') return true;
    return false;
}

// Same arithmetic as the reference's text->float conversion (paramdict.cpp:166-238): integer part,
// fraction as v2/10^n in double, power-of-ten scaling in double, one final rounding to float.
static float parse_float(const std::string& s) {
    const char* p = s.c_str();
    bool neg = *p == '-';
    if (*p == '+' || *p == '-') p++;
    unsigned int ip = 0;
    while (isdigit((unsigned char)*p)) ip = ip * 10 + (unsigned)(*p++ - '0');
    double v = (double)ip;
    if (*p == '.') {
        p++;
        unsigned int fp = 0, p10 = 1;
        while (isdigit((unsigned char)*p)) {
            fp = fp * 10 + (unsigned)(*p++ - '0');
            p10 *= 10;
        }
        v += fp / (double)p10;
    }
    if (*p == 'e' || *p == 'E') {
        p++;
        bool pos = *p != '-';
        if (*p == '+' || *p == '-') p++;
        unsigned int ex = 0;
        while (isdigit((unsigned char)*p)) ex = ex * 10 + (unsigned)(*p++ - '0');
        double sc = 1.0;
        while (ex >= 8) { sc *= 1e8; ex -= 8; }
        while (ex > 0) { sc *= 10.0; ex--; }
        v = pos ? v * sc : v / sc;
    }
    return neg ? (float)-v : (float)v;
}

static float half_to_float(uint16_t h) {
    uint32_t sign = (uint32_t)(h >> 15) << 31, exp = (h >> 10) & 31, man = h & 1023, bits;
    if (exp == 0) {
        if (man == 0) bits = sign;
        else {
            int e = -1;
            do { e++; man <<= 1; } while (!(man & 1024));
            bits = sign | ((uint32_t)(127 - 15 - e) << 23) | ((man & 1023) << 13);
        }
    } else if (exp == 31) bits = sign | 0x7F800000u | (man << 13);
    else bits = sign | ((exp + 112) << 23) | (man << 13);
    float f;
    memcpy(&f, &bits, 4);
    return f;
}

struct BinReader {
    const std::vector<unsigned char>& d;
    size_t pos = 0;
    bool ok = true;
    explicit BinReader(const std::vector<unsigned char>& v) : d(v) {}
    bool take(void* dst, size_t n) {
        if (pos + n > d.size()) { ok = false; return false; }
        memcpy(dst, d.data() + pos, n);
        pos += n;
        return true;
    }
    // tagged blob (conv / deconv / innerproduct weights)
    bool tagged(size_t n, std::vector<float>& out, bool& fp16) {
        uint32_t tag;
        if (!take(&tag, 4)) return false;
        if (n > d.size()) { ok = false; return false; }  // a count from a damaged .param must not drive the allocation below
        out.resize(n);
        if (tag == 0x01306B47u) {  // fp16 payload padded to 4 bytes
            fp16 = true;
            size_t bytes = (n * 2 + 3) & ~(size_t)3;
            if (pos + bytes > d.size()) { ok = false; return false; }
            for (size_t i = 0; i < n; i++) {
                uint16_t h;
                memcpy(&h, d.data() + pos + 2 * i, 2);
                out[i] = half_to_float(h);
            }
            pos += bytes;
            return true;
        }
        if (tag == 0) {  // raw fp32
            fp16 = false;
            return take(out.data(), n * 4);
        }
        ok = false;  // int8 / quantised tables are not used by any RIFE model
        return false;
    }
    bool raw(size_t n, std::vector<float>& out) {
        if (n > d.size() / 4 || pos + n * 4 > d.size()) { ok = false; return false; }
        out.resize(n);
        return n == 0 || take(out.data(), n * 4);
    }
};

static bool slurp(const std::string& path, std::string& out) {
    std::ifstream f(path, std::ios::binary);
    if (!f) return false;
    out.assign((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    return true;
}

int load_net(const std::string& param_path, const std::string& bin_path, Net& net, std::string& err) {
    std::string ptxt, bbytes;
    if (!slurp(param_path, ptxt)) { err = "cannot open " + param_path; return -1; }
    if (!slurp(bin_path, bbytes)) { err = "cannot open " + bin_path; return -1; }
    return parse_net(ptxt, bbytes, param_path, net, err);
}

int parse_net(const std::string& param_text, const std::string& bin_bytes, const std::string& param_path, Net& net, std::string& err) {
    const std::string& bin_path = param_path;
    std::istringstream pf(param_text);
    std::string line;
    int magic = 0, nl = 0, nb = 0;
    if (!(pf >> magic) || magic != 7767517) { err = "bad magic in " + param_path; return -2; }
    if (!(pf >> nl >> nb) || nl <= 0 || nb <= 0) { err = "bad counts in " + param_path; return -2; }
    std::getline(pf, line);
    net.layers.clear();
    net.blob_names.clear();
    std::map<std::string, int> blob_id;
    auto blob = [&](const std::string& n) {
        auto it = blob_id.find(n);
        if (it != blob_id.end()) return it->second;
        int id = (int)net.blob_names.size();
        blob_id[n] = id;
        net.blob_names.push_back(n);
        return id;
    };
    while (std::getline(pf, line)) {
        std::istringstream ss(line);
        Layer L;
        int nbot = 0, ntop = 0;
        if (!(ss >> L.type >> L.name >> nbot >> ntop)) continue;
        if (nbot < 0 || ntop < 0 || nbot > 4096 || ntop > 4096) { err = "bad blob counts at layer " + L.name + " in " + param_path; return -2; }
        std::string tok;
        for (int i = 0; i < nbot; i++) {
            if (!(ss >> tok)) { err = "missing bottom blob at layer " + L.name + " in " + param_path; return -2; }
            L.bottoms.push_back(blob(tok));
        }
        for (int i = 0; i < ntop; i++) {
            if (!(ss >> tok)) { err = "missing top blob at layer " + L.name + " in " + param_path; return -2; }
            L.tops.push_back(blob(tok));
        }
        while (ss >> tok) {
            size_t eq = tok.find('=');
            if (eq == std::string::npos) continue;
            int id = atoi(tok.substr(0, eq).c_str());
            std::string val = tok.substr(eq + 1);
            ParamVal pv;
            if (id <= -23300) {
                id = -id - 23300;
                pv.is_array = true;
                std::vector<std::string> parts;
                std::string cur;
                for (char c : val) {
                    if (c == ',') { parts.push_back(cur); cur.clear(); }
                    else cur += c;
                }
                parts.push_back(cur);
                size_t n = (size_t)atoi(parts[0].c_str());
                for (size_t k = 1; k < parts.size() && k <= n; k++) {
                    bool fl = token_is_float(parts[k]);
                    pv.is_float = pv.is_float || fl;
                    float f = fl ? parse_float(parts[k]) : (float)atoi(parts[k].c_str());
                    pv.af.push_back(f);
                    pv.ai.push_back(fl ? (int)f : atoi(parts[k].c_str()));
                }
            } else if (token_is_float(val)) {
                pv.is_float = true;
                pv.f = parse_float(val);
            } else {
                pv.i = atoi(val.c_str());
            }
            L.params[id] = pv;
        }
        net.layers.push_back(L);
    }
    if ((int)net.layers.size() != nl) { err = "layer count mismatch in " + param_path; return -2; }
    if ((int)net.blob_names.size() > nb) { err = "more blobs than declared in " + param_path; return -2; }
    net.producer.assign(net.blob_names.size(), -1);
    for (size_t li = 0; li < net.layers.size(); li++)
        for (int t : net.layers[li].tops) net.producer[t] = (int)li;

    std::vector<unsigned char> bytes(bin_bytes.begin(), bin_bytes.end());
    BinReader br(bytes);
    for (Layer& L : net.layers) {
        if (L.type == "Convolution" || L.type == "Deconvolution" || L.type == "InnerProduct") {
            int num_output = L.geti(0, 0);
            bool has_bias = L.type == "InnerProduct" ? L.geti(1, 0) != 0 : L.geti(5, 0) != 0;
            int wsize = L.type == "InnerProduct" ? L.geti(2, 0) : L.geti(6, 0);
            if (wsize < 0 || num_output < 0) br.ok = false;
            else {
                br.tagged((size_t)wsize, L.weight, L.weight_is_fp16);
                if (br.ok && has_bias) br.raw((size_t)num_output, L.bias);
            }
        } else if (L.type == "PReLU") {
            int ns = L.geti(0, 0);
            if (ns < 0) br.ok = false;
            else br.raw((size_t)ns, L.slope);
        }
        if (!br.ok) { err = "truncated or unsupported weights in " + bin_path + " at layer " + L.name; return -3; }
    }
    if (br.pos != bytes.size()) { err = "trailing bytes in " + bin_path; return -3; }
    return 0;
}

}  // namespace rife
