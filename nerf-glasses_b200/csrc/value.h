// value.h - a small dynamic value tree shared by the msgpack reader (iNGP snapshots) and the JSON reader (glTF).
// Replaces the reference's use of nlohmann::json (R/dependencies/json) for the two formats the hot path loads.
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace nmr {

struct Value {
    enum Type { Null, Bool, Int, Float, String, Binary, Array, Map } type = Null;
    bool b = false;
    int64_t i = 0;
    double f = 0.0;
    std::string s;                       // String payload
    const uint8_t* bin = nullptr;        // Binary payload: points into the file buffer (zero-copy)
    size_t bin_size = 0;
    std::vector<Value> arr;
    std::vector<std::pair<std::string, Value>> map;   // insertion order kept

    bool is_null() const { return type == Null; }
    bool is_number() const { return type == Int || type == Float; }
    bool is_array() const { return type == Array; }
    bool is_map() const { return type == Map; }

    const Value* find(const std::string& key) const {
        if (type != Map) return nullptr;
        for (const auto& kv : map) if (kv.first == key) return &kv.second;
        return nullptr;
    }
    bool contains(const std::string& key) const { return find(key) != nullptr; }
    const Value& at(const std::string& key) const {
        const Value* v = find(key);
        if (!v) throw std::runtime_error("missing key '" + key + "'");
        return *v;
    }
    const Value& at(size_t idx) const {
        if (type != Array || idx >= arr.size()) throw std::runtime_error("array index out of range");
        return arr[idx];
    }
    size_t size() const { return type == Array ? arr.size() : (type == Map ? map.size() : 0); }

    double as_double() const {
        if (type == Float) return f;
        if (type == Int) return (double)i;
        if (type == Bool) return b ? 1.0 : 0.0;
        throw std::runtime_error("value is not a number");
    }
    float as_float() const { return (float)as_double(); }
    int64_t as_int() const {
        if (type == Int) return i;
        if (type == Float) return (int64_t)f;
        if (type == Bool) return b ? 1 : 0;
        throw std::runtime_error("value is not an integer");
    }
    bool as_bool() const {
        if (type == Bool) return b;
        if (type == Int) return i != 0;
        throw std::runtime_error("value is not a bool");
    }
    const std::string& as_string() const {
        if (type != String) throw std::runtime_error("value is not a string");
        return s;
    }
    double value(const std::string& key, double dflt) const { const Value* v = find(key); return (v && v->is_number()) ? v->as_double() : dflt; }
    int64_t value_int(const std::string& key, int64_t dflt) const { const Value* v = find(key); return (v && (v->is_number() || v->type == Bool)) ? v->as_int() : dflt; }
    std::string value_str(const std::string& key, const std::string& dflt) const { const Value* v = find(key); return (v && v->type == String) ? v->s : dflt; }
    bool value_bool(const std::string& key, bool dflt) const { const Value* v = find(key); return (v && (v->type == Bool || v->type == Int)) ? v->as_bool() : dflt; }
};

// Parses one msgpack object from [data, data+size).  Binary payloads alias `data`.
Value parse_msgpack(const uint8_t* data, size_t size);
// Parses a JSON document.
Value parse_json(const char* text, size_t size);

std::vector<uint8_t> read_file(const std::string& path);   // throws std::runtime_error("cannot open ...")

}  // namespace nmr
