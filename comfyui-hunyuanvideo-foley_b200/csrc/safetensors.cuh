// safetensors -> engine layout, directly.  Replaces the reference Model Loader's
// load_torch_file -> init_empty_weights -> to_empty -> load_state_dict -> .to(dtype) -> optional FP8 wrap chain
// (nodes.py:85-126): the file is mapped, its JSON header parsed here, and every tensor the hot path uses goes from
// the mapping straight into device memory, where finalize() repacks it.  FP8 checkpoints (utils.py:492-503) are
// de-quantised on the device at repack time; `quantization != none` (the reference stores wrapped weights in FP8 and
// upcasts them per forward, utils.py:316-485) is honoured as a rounding of exactly those weights through the FP8
// format, so the numbers match the reference's while the GEMMs keep reading bf16.
#pragma once
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "common.cuh"

namespace foley {

struct StEntry {
    std::string name;
    std::string dtype;            // "BF16", "F32", "F16", "F8_E4M3", "F8_E5M2", ...
    std::vector<int64_t> shape;
    uint64_t begin = 0, end = 0;  // byte offsets into the data section
};

// Minimal JSON reader for the safetensors header: one object of {name: {"dtype": str, "shape": [int...],
// "data_offsets": [int, int]}} plus an optional "__metadata__" object of strings.  Anything else is a format error.
class StHeaderParser {
  public:
    StHeaderParser(const char* p, size_t n) : p_(p), end_(p + n) {}
    bool parse(std::vector<StEntry>* out, std::string* err) {
        ws();
        if (!eat('{')) return bad("header is not a JSON object", err);
        ws();
        if (peek() == '}') { ++p_; return true; }
        for (;;) {
            std::string key;
            ws();
            if (!str(&key)) return bad("expected a tensor name", err);
            ws();
            if (!eat(':')) return bad("expected ':' after " + key, err);
            ws();
            if (key == "__metadata__") {
                if (!skip_value()) return bad("malformed __metadata__", err);
            } else {
                StEntry e;
                e.name = key;
                if (!entry(&e)) return bad("malformed entry for " + key, err);
                out->push_back(std::move(e));
            }
            ws();
            if (eat(',')) continue;
            if (eat('}')) return true;
            return bad("expected ',' or '}' after " + key, err);
        }
    }

  private:
    const char* p_;
    const char* end_;
    char peek() const { return p_ < end_ ? *p_ : '\0'; }
    bool eat(char c) { if (peek() == c) { ++p_; return true; } return false; }
    void ws() { while (p_ < end_ && (*p_ == ' ' || *p_ == '\n' || *p_ == '\t' || *p_ == '\r')) ++p_; }
    static bool bad(const std::string& m, std::string* err) { if (err) *err = "safetensors: " + m; return false; }
    bool str(std::string* out) {
        if (!eat('"')) return false;
        out->clear();
        while (p_ < end_ && *p_ != '"') {
            if (*p_ == '\\') {
                if (++p_ >= end_) return false;
                switch (*p_) {
                    case 'n': out->push_back('\n'); break;
                    case 't': out->push_back('\t'); break;
                    case 'r': out->push_back('\r'); break;
                    case 'b': out->push_back('\b'); break;
                    case 'f': out->push_back('\f'); break;
                    case 'u': {   // tensor names are ASCII in practice; keep the low byte of the code unit
                        if (end_ - p_ < 5) return false;
                        unsigned v = 0;
                        for (int i = 1; i <= 4; ++i) {
                            const char c = p_[i];
                            v = v * 16 + (c >= '0' && c <= '9' ? c - '0' : (c | 32) >= 'a' && (c | 32) <= 'f' ? (c | 32) - 'a' + 10 : 0);
                        }
                        out->push_back(static_cast<char>(v & 0x7F));
                        p_ += 4;
                        break;
                    }
                    default: out->push_back(*p_);   // \" \\ \/
                }
                ++p_;
            } else {
                out->push_back(*p_++);
            }
        }
        return eat('"');
    }
    bool uint(uint64_t* v) {
        if (p_ >= end_ || *p_ < '0' || *p_ > '9') return false;
        uint64_t x = 0;
        while (p_ < end_ && *p_ >= '0' && *p_ <= '9') x = x * 10 + static_cast<uint64_t>(*p_++ - '0');
        *v = x;
        return true;
    }
    bool uint_array(std::vector<uint64_t>* out) {
        if (!eat('[')) return false;
        ws();
        if (eat(']')) return true;
        for (;;) {
            uint64_t v;
            ws();
            if (!uint(&v)) return false;
            out->push_back(v);
            ws();
            if (eat(',')) continue;
            return eat(']');
        }
    }
    static constexpr int kMaxDepth = 64;   // nesting limit of skipped values: recursion depth must not follow the input
    bool skip_value(int depth = 0) {   // strings, numbers, literals, nested arrays / objects
        ws();
        const char c = peek();
        if (c == '"') { std::string s; return str(&s); }
        if (c == '{' || c == '[') {
            if (depth >= kMaxDepth) return false;
            const char close = c == '{' ? '}' : ']';
            ++p_;
            ws();
            if (eat(close)) return true;
            for (;;) {
                if (c == '{') {
                    std::string k;
                    ws();
                    if (!str(&k)) return false;
                    ws();
                    if (!eat(':')) return false;
                }
                if (!skip_value(depth + 1)) return false;
                ws();
                if (eat(',')) continue;
                return eat(close);
            }
        }
        const char* s = p_;
        while (p_ < end_ && *p_ != ',' && *p_ != '}' && *p_ != ']' && *p_ != ' ' && *p_ != '\n') ++p_;
        return p_ > s;
    }
    bool entry(StEntry* e) {
        if (!eat('{')) return false;
        bool have_dtype = false, have_shape = false, have_off = false;
        for (;;) {
            std::string k;
            ws();
            if (!str(&k)) return false;
            ws();
            if (!eat(':')) return false;
            ws();
            if (k == "dtype") { if (!str(&e->dtype)) return false; have_dtype = true; }
            else if (k == "shape") {
                std::vector<uint64_t> v;
                if (!uint_array(&v)) return false;
                for (uint64_t d : v) e->shape.push_back(static_cast<int64_t>(d));
                have_shape = true;
            } else if (k == "data_offsets") {
                std::vector<uint64_t> v;
                if (!uint_array(&v) || v.size() != 2 || v[1] < v[0]) return false;
                e->begin = v[0]; e->end = v[1];
                have_off = true;
            } else if (!skip_value()) return false;
            ws();
            if (eat(',')) continue;
            if (eat('}')) return have_dtype && have_shape && have_off;
            return false;
        }
    }
};

inline int st_dtype_bytes(const std::string& d) {
    if (d == "F64" || d == "I64" || d == "U64") return 8;
    if (d == "F32" || d == "I32" || d == "U32") return 4;
    if (d == "BF16" || d == "F16" || d == "I16" || d == "U16") return 2;
    if (d == "F8_E4M3" || d == "F8_E5M2" || d == "I8" || d == "U8" || d == "BOOL") return 1;
    return 0;
}
inline int st_dtype_to_foley(const std::string& d) {
    if (d == "BF16") return FOLEY_DT_BF16;
    if (d == "F32") return FOLEY_DT_F32;
    if (d == "F16") return FOLEY_DT_F16;
    if (d == "F8_E4M3") return FOLEY_DT_F8_E4M3FN;
    if (d == "F8_E5M2") return FOLEY_DT_F8_E5M2;
    return -1;
}

// A read-only mapping of a .safetensors file with its parsed, validated header.
class StFile {
  public:
    std::vector<StEntry> entries;
    const uint8_t* data = nullptr;     // start of the data section
    uint64_t data_bytes = 0;
    ~StFile() { close_file(); }
    bool open_file(const char* path, std::string* err) {
        close_file();
        fd_ = ::open(path, O_RDONLY);
        if (fd_ < 0) return bad(std::string("cannot open ") + path, err);
        struct stat st;
        if (fstat(fd_, &st) != 0 || st.st_size < 8) return bad(std::string("not a safetensors file: ") + path, err);
        size_ = static_cast<size_t>(st.st_size);
        map_ = mmap(nullptr, size_, PROT_READ, MAP_PRIVATE, fd_, 0);
        if (map_ == MAP_FAILED) { map_ = nullptr; return bad(std::string("mmap failed: ") + path, err); }
        const uint8_t* b = static_cast<const uint8_t*>(map_);
        uint64_t hlen = 0;
        for (int i = 7; i >= 0; --i) hlen = (hlen << 8) | b[i];   // little-endian u64
        if (hlen > size_ - 8 || hlen > (100ull << 20)) return bad("header length out of range", err);
        StHeaderParser parser(reinterpret_cast<const char*>(b + 8), static_cast<size_t>(hlen));
        if (!parser.parse(&entries, err)) return false;
        data = b + 8 + hlen;
        data_bytes = size_ - 8 - hlen;
        for (const StEntry& e : entries) {
            const int eb = st_dtype_bytes(e.dtype);
            if (eb == 0) return bad("unknown dtype " + e.dtype + " for " + e.name, err);
            uint64_t numel = 1;
            bool overflow = false;
            for (int64_t d : e.shape) {
                if (d < 0 || (d != 0 && numel > (1ull << 46) / static_cast<uint64_t>(d))) { overflow = true; break; }
                numel *= static_cast<uint64_t>(d);
            }
            if (overflow) return bad("shape of " + e.name + " is out of range", err);
            if (e.end > data_bytes || e.end - e.begin != numel * static_cast<uint64_t>(eb))
                return bad("data_offsets do not match shape/dtype for " + e.name, err);
        }
        return true;
    }
    void close_file() {
        if (map_) munmap(map_, size_);
        if (fd_ >= 0) ::close(fd_);
        map_ = nullptr; fd_ = -1; size_ = 0; entries.clear(); data = nullptr; data_bytes = 0;
    }

  private:
    int fd_ = -1;
    void* map_ = nullptr;
    size_t size_ = 0;
    static bool bad(const std::string& m, std::string* err) { if (err) *err = "safetensors: " + m; return false; }
};

// Which weights the reference's FP8 wrap replaces (utils.py:298-311, 433-481): nominally every nn.Linear / nn.Conv1d
// whose qualified module name contains none of the deny substrings.  The recursion builds that name as
// f"{prefix}{name}" — WITHOUT a dot between the levels ("triple_blocks0audio_mlpfc1") — while every deny token
// contains a dot, so nothing is ever denied: running the reference's _wrap_fp8_inplace on its own model wraps every
// Linear / Conv1d, final_layer.* and visual_proj.* included (probe: tools/make_golden.py fp8_wrap, 56 of 56 modules on
// the tiny config; tests/golden/fp8_wrapped_tiny.json).  Reproduced as is.  Linear / Conv weights are the >= 2-D
// "*.weight" tensors (norm weights are 1-D; the 2-D sync_pos_emb / empty_*_feat are bare parameters, not modules).
inline bool fp8_wraps(const std::string& tensor_name, int ndim) {
    const std::string suffix = ".weight";
    if (ndim < 2 || tensor_name.size() <= suffix.size() ||
        tensor_name.compare(tensor_name.size() - suffix.size(), suffix.size(), suffix) != 0)
        return false;
    if (tensor_name.rfind("dac.", 0) == 0) return false;   // the DAC-VAE is a separate, never-wrapped model
    return true;
}

}  // namespace foley
