// Minimal JSON reader for the setup blob (serde_json of StarkInfo / Program / StarkStruct).  Header-only.
#pragma once
#include <string>
#include <vector>
#include <map>
#include <memory>
#include <stdexcept>
#include <cstdlib>
#include <cstring>

namespace mj {
struct Value;
typedef std::shared_ptr<Value> P;
struct Value {
    enum Kind { Null, Bool, Num, Str, Arr, Obj } kind = Null;
    bool b = false; double num = 0; long long inum = 0; std::string str;
    std::vector<P> arr; std::vector<std::pair<std::string, P>> obj;
    const Value& at(const std::string& k) const {
        for (auto& kv : obj) if (kv.first == k) return *kv.second;
        throw std::runtime_error("json: missing key '" + k + "'");
    }
    bool has(const std::string& k) const { for (auto& kv : obj) if (kv.first == k) return true; return false; }
    const Value& operator[](size_t i) const { if (i >= arr.size()) throw std::runtime_error("json: index out of range"); return *arr[i]; }
    size_t size() const { return kind == Arr ? arr.size() : obj.size(); }
    long long as_int() const { if (kind != Num) throw std::runtime_error("json: not a number"); return inum; }
    size_t as_size() const { return (size_t)as_int(); }
    bool as_bool() const { if (kind != Bool) throw std::runtime_error("json: not a bool"); return b; }
    const std::string& as_str() const { if (kind != Str) throw std::runtime_error("json: not a string"); return str; }
    bool is_null() const { return kind == Null; }
};
class Parser {
    const char* p; const char* e;
    void ws() { while (p < e && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) p++; }
    [[noreturn]] void fail(const char* m) { throw std::runtime_error(std::string("json parse error: ") + m); }
    std::string str() {
        if (*p != '"') fail("expected string"); p++;
        std::string s;
        while (p < e && *p != '"') {
            if (*p == '\\') { p++; if (p >= e) fail("bad escape");
                switch (*p) { case 'n': s += '\n'; break; case 't': s += '\t'; break; case 'r': s += '\r'; break; case 'b': s += '\b'; break; case 'f': s += '\f'; break;
                    case 'u': { if (e - p < 5) fail("bad \\u"); unsigned c = (unsigned)strtoul(std::string(p + 1, p + 5).c_str(), nullptr, 16); s += (char)(c < 128 ? c : '?'); p += 4; break; }
                    default: s += *p; }
                p++; }
            else s += *p++;
        }
        if (p >= e) fail("unterminated string"); p++;
        return s;
    }
    P val() {
        ws(); if (p >= e) fail("eof");
        P v = std::make_shared<Value>();
        if (*p == '{') { p++; v->kind = Value::Obj; ws(); if (*p == '}') { p++; return v; }
            for (;;) { ws(); std::string k = str(); ws(); if (*p != ':') fail("expected :"); p++; P c = val(); v->obj.emplace_back(k, c); ws();
                if (*p == ',') { p++; continue; } if (*p == '}') { p++; break; } fail("expected , or }"); } }
        else if (*p == '[') { p++; v->kind = Value::Arr; ws(); if (*p == ']') { p++; return v; }
            for (;;) { v->arr.push_back(val()); ws(); if (*p == ',') { p++; continue; } if (*p == ']') { p++; break; } fail("expected , or ]"); } }
        else if (*p == '"') { v->kind = Value::Str; v->str = str(); }
        else if (!strncmp(p, "true", 4)) { v->kind = Value::Bool; v->b = true; p += 4; }
        else if (!strncmp(p, "false", 5)) { v->kind = Value::Bool; v->b = false; p += 5; }
        else if (!strncmp(p, "null", 4)) { v->kind = Value::Null; p += 4; }
        else { char* end; v->kind = Value::Num; v->inum = strtoll(p, &end, 10); v->num = (double)v->inum;
            if (end == p) fail("bad token"); if (end < e && (*end == '.' || *end == 'e' || *end == 'E')) { v->num = strtod(p, &end); v->inum = (long long)v->num; } p = end; }
        return v;
    }
public:
    static P parse(const std::string& s) { Parser q; q.p = s.data(); q.e = s.data() + s.size(); P v = q.val(); q.ws(); if (q.p != q.e) q.fail("trailing data"); return v; }
};
}  // namespace mj
