// value.cpp - msgpack and JSON readers (see value.h).
// msgpack: the container format of iNGP snapshots, written by nlohmann::json::to_msgpack upstream and read by
// json::from_msgpack in the reference (S/ngp/testbed.cu:1019-1020).  Supports the whole core spec except ext types'
// interpretation (they are skipped into Null) - snapshots only use map/array/str/bin/float/int/bool/nil.
#include "value.h"

#include <cmath>
#include <cstdio>
#include <cstring>

namespace nmr {

std::vector<uint8_t> read_file(const std::string& path) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("cannot open '" + path + "'");
    std::fseek(f, 0, SEEK_END);
    long n = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    if (n < 0) { std::fclose(f); throw std::runtime_error("cannot stat '" + path + "'"); }
    std::vector<uint8_t> buf((size_t)n);
    size_t got = n ? std::fread(buf.data(), 1, (size_t)n, f) : 0;
    std::fclose(f);
    if (got != (size_t)n) throw std::runtime_error("short read on '" + path + "'");
    return buf;
}

namespace {

struct MsgpackReader {
    const uint8_t* p;
    const uint8_t* end;
    int depth = 0;

    void need(size_t n) const { if ((size_t)(end - p) < n) throw std::runtime_error("msgpack: truncated input"); }
    uint8_t u8() { need(1); return *p++; }
    uint16_t u16() { need(2); uint16_t v = (uint16_t)((p[0] << 8) | p[1]); p += 2; return v; }
    uint32_t u32() { need(4); uint32_t v = ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; p += 4; return v; }
    uint64_t u64() { uint64_t hi = u32(); uint64_t lo = u32(); return (hi << 32) | lo; }

    Value str(size_t n) { need(n); Value v; v.type = Value::String; v.s.assign((const char*)p, n); p += n; return v; }
    Value bin(size_t n) { need(n); Value v; v.type = Value::Binary; v.bin = p; v.bin_size = n; p += n; return v; }
    Value array(size_t n) {
        Value v; v.type = Value::Array; v.arr.reserve(n < 65536 ? n : 65536);
        for (size_t k = 0; k < n; ++k) v.arr.push_back(parse());
        return v;
    }
    Value map(size_t n) {
        Value v; v.type = Value::Map; v.map.reserve(n < 4096 ? n : 4096);
        for (size_t k = 0; k < n; ++k) {
            Value key = parse();
            std::string ks;
            if (key.type == Value::String) ks = key.s;
            else if (key.type == Value::Int) ks = std::to_string(key.i);
            else throw std::runtime_error("msgpack: unsupported map key type");
            v.map.emplace_back(std::move(ks), parse());
        }
        return v;
    }
    Value integer(int64_t x) { Value v; v.type = Value::Int; v.i = x; return v; }
    Value real(double x) { Value v; v.type = Value::Float; v.f = x; return v; }

    Value parse() {
        if (++depth > 256) throw std::runtime_error("msgpack: nesting too deep");
        struct Guard { int& d; ~Guard() { --d; } } g{depth};
        uint8_t t = u8();
        if (t <= 0x7f) return integer(t);
        if (t >= 0xe0) return integer((int8_t)t);
        if (t >= 0x80 && t <= 0x8f) return map(t & 0x0f);
        if (t >= 0x90 && t <= 0x9f) return array(t & 0x0f);
        if (t >= 0xa0 && t <= 0xbf) return str(t & 0x1f);
        switch (t) {
            case 0xc0: return Value{};
            case 0xc2: { Value v; v.type = Value::Bool; v.b = false; return v; }
            case 0xc3: { Value v; v.type = Value::Bool; v.b = true; return v; }
            case 0xc4: return bin(u8());
            case 0xc5: return bin(u16());
            case 0xc6: return bin(u32());
            case 0xc7: { size_t n = u8(); u8(); need(n); p += n; return Value{}; }
            case 0xc8: { size_t n = u16(); u8(); need(n); p += n; return Value{}; }
            case 0xc9: { size_t n = u32(); u8(); need(n); p += n; return Value{}; }
            case 0xca: { uint32_t b = u32(); float f; std::memcpy(&f, &b, 4); return real(f); }
            case 0xcb: { uint64_t b = u64(); double d; std::memcpy(&d, &b, 8); return real(d); }
            case 0xcc: return integer(u8());
            case 0xcd: return integer(u16());
            case 0xce: return integer(u32());
            case 0xcf: return integer((int64_t)u64());
            case 0xd0: return integer((int8_t)u8());
            case 0xd1: return integer((int16_t)u16());
            case 0xd2: return integer((int32_t)u32());
            case 0xd3: return integer((int64_t)u64());
            case 0xd4: need(2); p += 2; return Value{};
            case 0xd5: need(3); p += 3; return Value{};
            case 0xd6: need(5); p += 5; return Value{};
            case 0xd7: need(9); p += 9; return Value{};
            case 0xd8: need(17); p += 17; return Value{};
            case 0xd9: return str(u8());
            case 0xda: return str(u16());
            case 0xdb: return str(u32());
            case 0xdc: return array(u16());
            case 0xdd: return array(u32());
            case 0xde: return map(u16());
            case 0xdf: return map(u32());
            default: throw std::runtime_error("msgpack: reserved type byte 0xc1");
        }
    }
};

struct JsonReader {
    const char* p;
    const char* end;
    int depth = 0;

    [[noreturn]] void fail(const char* what) const { throw std::runtime_error(std::string("json: ") + what); }
    void ws() { while (p < end && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) ++p; }
    bool eat(char c) { ws(); if (p < end && *p == c) { ++p; return true; } return false; }
    void expect(char c) { if (!eat(c)) fail("unexpected character"); }

    static void append_utf8(std::string& out, uint32_t cp) {
        if (cp < 0x80) out.push_back((char)cp);
        else if (cp < 0x800) { out.push_back((char)(0xC0 | (cp >> 6))); out.push_back((char)(0x80 | (cp & 0x3F))); }
        else if (cp < 0x10000) { out.push_back((char)(0xE0 | (cp >> 12))); out.push_back((char)(0x80 | ((cp >> 6) & 0x3F))); out.push_back((char)(0x80 | (cp & 0x3F))); }
        else { out.push_back((char)(0xF0 | (cp >> 18))); out.push_back((char)(0x80 | ((cp >> 12) & 0x3F))); out.push_back((char)(0x80 | ((cp >> 6) & 0x3F))); out.push_back((char)(0x80 | (cp & 0x3F))); }
    }
    uint32_t hex4() {
        if (end - p < 4) fail("truncated \\u escape");
        uint32_t v = 0;
        for (int k = 0; k < 4; ++k) {
            char c = *p++; v <<= 4;
            if (c >= '0' && c <= '9') v |= (uint32_t)(c - '0');
            else if (c >= 'a' && c <= 'f') v |= (uint32_t)(c - 'a' + 10);
            else if (c >= 'A' && c <= 'F') v |= (uint32_t)(c - 'A' + 10);
            else fail("bad \\u escape");
        }
        return v;
    }
    std::string string_body() {
        std::string out;
        while (true) {
            if (p >= end) fail("unterminated string");
            char c = *p++;
            if (c == '"') break;
            if (c == '\\') {
                if (p >= end) fail("unterminated escape");
                char e = *p++;
                switch (e) {
                    case '"': out.push_back('"'); break; case '\\': out.push_back('\\'); break; case '/': out.push_back('/'); break;
                    case 'b': out.push_back('\b'); break; case 'f': out.push_back('\f'); break; case 'n': out.push_back('\n'); break;
                    case 'r': out.push_back('\r'); break; case 't': out.push_back('\t'); break;
                    case 'u': {
                        uint32_t cp = hex4();
                        if (cp >= 0xD800 && cp <= 0xDBFF && end - p >= 6 && p[0] == '\\' && p[1] == 'u') { p += 2; uint32_t lo = hex4(); cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00); }
                        append_utf8(out, cp); break;
                    }
                    default: fail("bad escape");
                }
            } else out.push_back(c);
        }
        return out;
    }
    Value parse() {
        if (++depth > 256) fail("nesting too deep");
        struct Guard { int& d; ~Guard() { --d; } } g{depth};
        ws();
        if (p >= end) fail("unexpected end");
        char c = *p;
        if (c == '{') {
            ++p; Value v; v.type = Value::Map;
            if (eat('}')) return v;
            do { ws(); expect('"'); std::string k = string_body(); expect(':'); v.map.emplace_back(std::move(k), parse()); } while (eat(','));
            expect('}'); return v;
        }
        if (c == '[') {
            ++p; Value v; v.type = Value::Array;
            if (eat(']')) return v;
            do { v.arr.push_back(parse()); } while (eat(','));
            expect(']'); return v;
        }
        if (c == '"') { ++p; Value v; v.type = Value::String; v.s = string_body(); return v; }
        if (end - p >= 4 && !std::strncmp(p, "true", 4)) { p += 4; Value v; v.type = Value::Bool; v.b = true; return v; }
        if (end - p >= 5 && !std::strncmp(p, "false", 5)) { p += 5; Value v; v.type = Value::Bool; v.b = false; return v; }
        if (end - p >= 4 && !std::strncmp(p, "null", 4)) { p += 4; return Value{}; }
        const char* s = p; bool is_float = false;
        if (p < end && (*p == '-' || *p == '+')) ++p;
        while (p < end && ((*p >= '0' && *p <= '9') || *p == '.' || *p == 'e' || *p == 'E' || *p == '-' || *p == '+')) { if (*p == '.' || *p == 'e' || *p == 'E') is_float = true; ++p; }
        if (p == s) fail("unexpected token");
        std::string num(s, p);
        Value v;
        if (is_float) { v.type = Value::Float; v.f = std::strtod(num.c_str(), nullptr); }
        else { v.type = Value::Int; v.i = std::strtoll(num.c_str(), nullptr, 10); }
        return v;
    }
};

}  // namespace

Value parse_msgpack(const uint8_t* data, size_t size) {
    MsgpackReader r{data, data + size};
    return r.parse();
}

Value parse_json(const char* text, size_t size) {
    JsonReader r{text, text + size};
    Value v = r.parse();
    r.ws();
    if (r.p != r.end) throw std::runtime_error("json: trailing characters");
    return v;
}

}  // namespace nmr
