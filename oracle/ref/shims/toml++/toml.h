// Build shim (test infrastructure, see oracle/README.md).
//
// The reference pulls toml++ 3.3.0 through CMake FetchContent, which is unreachable in
// this sandbox.  This header is a from-scratch stand-in that provides the slice of the
// toml++ API the reference headers name (node / value<T> / table / array / node_view /
// date_time / parse / parse_file / operator<<) plus a small parser for the TOML subset
// the reference's own files use (tables, arrays of tables, dotted keys, strings, ints,
// floats, bools, local date-times, single-line and multi-line arrays).
//
// It exists only so the reference's CPU path can be compiled as the parity oracle and
// so the host-side adapter can read `vamana_config.toml`.  It is not a general TOML
// implementation.
#pragma once

#include <cctype>
#include <cstdint>
#include <cstdlib>
#include <fstream>
#include <initializer_list>
#include <istream>
#include <map>
#include <memory>
#include <ostream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <string_view>
#include <type_traits>
#include <utility>
#include <vector>

namespace toml {

struct source_position {
    uint32_t line = 0;
    uint32_t column = 0;
};
struct source_region {
    source_position begin{};
    source_position end{};
};
inline std::ostream& operator<<(std::ostream& os, const source_region& r) {
    return os << "line " << r.begin.line;
}

struct date {
    int year = 0;
    unsigned month = 0;
    unsigned day = 0;
    date() = default;
    date(int y, unsigned m, unsigned d)
        : year(y)
        , month(m)
        , day(d) {}
};
struct time {
    uint8_t hour = 0;
    uint8_t minute = 0;
    uint8_t second = 0;
    uint32_t nanosecond = 0;
};
struct date_time {
    toml::date date{};
    toml::time time{};
    date_time() = default;
    date_time(toml::date d, toml::time t)
        : date(d)
        , time(t) {}
};

class table;
class array;
template <typename T> class value;

class node {
  public:
    virtual ~node() = default;
    virtual std::unique_ptr<node> clone() const = 0;
    virtual void print(std::ostream& os, int indent) const = 0;
    const source_region& source() const { return src_; }
    void set_line(uint32_t line) { src_.begin.line = line; }

    template <typename T> const T* as() const { return dynamic_cast<const T*>(this); }
    template <typename T> T* as() { return dynamic_cast<T*>(this); }
    const table* as_table() const;
    const array* as_array() const;
    bool is_table() const { return as_table() != nullptr; }
    bool is_array() const { return as_array() != nullptr; }

    template <typename F> decltype(auto) visit(F&& f) const;

  private:
    source_region src_{};
};

namespace impl {
inline void print_scalar(std::ostream& os, const std::string& v) { os << '\'' << v << '\''; }
inline void print_scalar(std::ostream& os, bool v) { os << (v ? "true" : "false"); }
inline void print_scalar(std::ostream& os, int64_t v) { os << v; }
inline void print_scalar(std::ostream& os, double v) {
    std::ostringstream s;
    s.precision(17);
    s << v;
    auto str = s.str();
    if (str.find_first_of(".eEn") == std::string::npos) {
        str += ".0";
    }
    os << str;
}
inline void print_scalar(std::ostream& os, const date_time& v) {
    char buf[40];
    std::snprintf(
        buf,
        sizeof(buf),
        "%04d-%02u-%02uT%02u:%02u:%02u",
        v.date.year,
        v.date.month,
        v.date.day,
        unsigned(v.time.hour),
        unsigned(v.time.minute),
        unsigned(v.time.second)
    );
    os << buf;
}
} // namespace impl

template <typename T> class value : public node {
  public:
    value() = default;
    value(T v)
        : v_(std::move(v)) {}
    const T& get() const { return v_; }
    T& get() { return v_; }
    const T& operator*() const { return v_; }
    std::unique_ptr<node> clone() const override {
        auto p = std::make_unique<value<T>>(v_);
        p->set_line(source().begin.line);
        return p;
    }
    void print(std::ostream& os, int) const override { impl::print_scalar(os, v_); }

  private:
    T v_{};
};

template <typename N> class node_view {
  public:
    node_view() = default;
    explicit node_view(N* n)
        : n_(n) {}
    explicit operator bool() const { return n_ != nullptr; }
    N* node() const { return n_; }
    const table* as_table() const { return n_ ? n_->as_table() : nullptr; }
    const array* as_array() const { return n_ ? n_->as_array() : nullptr; }
    template <typename T> auto as() const { return n_ ? n_->template as<T>() : nullptr; }

  private:
    N* n_ = nullptr;
};

namespace impl {
template <typename T> std::unique_ptr<node> make_node(T&& v);

struct table_init_pair {
    std::string key;
    std::unique_ptr<node> value;
    template <typename K, typename V>
    table_init_pair(K&& k, V&& v)
        : key(std::forward<K>(k))
        , value(make_node(std::forward<V>(v))) {}
};
} // namespace impl

class array : public node {
  public:
    array() = default;
    array(const array& o)
        : node(o) {
        for (const auto& e : o.elems_) {
            elems_.push_back(e->clone());
        }
    }
    array& operator=(const array& o) {
        if (this != &o) {
            elems_.clear();
            for (const auto& e : o.elems_) {
                elems_.push_back(e->clone());
            }
        }
        return *this;
    }
    array(array&&) = default;
    array& operator=(array&&) = default;

    std::unique_ptr<node> clone() const override { return std::make_unique<array>(*this); }
    void print(std::ostream& os, int indent) const override;

    template <typename T> void push_back(T&& v) {
        elems_.push_back(impl::make_node(std::forward<T>(v)));
    }
    template <typename T> void emplace_back_node(std::unique_ptr<T> p) {
        elems_.push_back(std::move(p));
    }
    size_t size() const { return elems_.size(); }
    bool empty() const { return elems_.empty(); }
    const node& operator[](size_t i) const { return *elems_[i]; }
    node& operator[](size_t i) { return *elems_[i]; }
    node* get(size_t i) { return i < elems_.size() ? elems_[i].get() : nullptr; }
    const node* get(size_t i) const { return i < elems_.size() ? elems_[i].get() : nullptr; }
    node& back() { return *elems_.back(); }

    template <typename Base, typename Ref> struct iter {
        Base it;
        Ref operator*() const { return **it; }
        iter& operator++() {
            ++it;
            return *this;
        }
        bool operator!=(const iter& o) const { return it != o.it; }
        bool operator==(const iter& o) const { return it == o.it; }
    };
    using storage = std::vector<std::unique_ptr<node>>;
    using const_iterator = iter<storage::const_iterator, const node&>;
    using iterator = iter<storage::iterator, node&>;
    const_iterator begin() const { return {elems_.begin()}; }
    const_iterator end() const { return {elems_.end()}; }
    iterator begin() { return {elems_.begin()}; }
    iterator end() { return {elems_.end()}; }

  private:
    storage elems_;
};

class table : public node {
  public:
    using map_type = std::map<std::string, std::unique_ptr<toml::node>, std::less<>>;

    table() = default;
    table(std::initializer_list<impl::table_init_pair> init) {
        for (const auto& p : init) {
            map_[p.key] = p.value->clone();
        }
    }
    table(const table& o)
        : node(o) {
        for (const auto& [k, v] : o.map_) {
            map_[k] = v->clone();
        }
    }
    table& operator=(const table& o) {
        if (this != &o) {
            map_.clear();
            for (const auto& [k, v] : o.map_) {
                map_[k] = v->clone();
            }
        }
        return *this;
    }
    table(table&&) = default;
    table& operator=(table&&) = default;

    std::unique_ptr<node> clone() const override { return std::make_unique<table>(*this); }
    void print(std::ostream& os, int indent) const override;

    node_view<const toml::node> operator[](std::string_view k) const {
        auto it = map_.find(k);
        return node_view<const toml::node>(it == map_.end() ? nullptr : it->second.get());
    }
    node_view<toml::node> operator[](std::string_view k) {
        auto it = map_.find(k);
        return node_view<toml::node>(it == map_.end() ? nullptr : it->second.get());
    }
    toml::node* get(std::string_view k) {
        auto it = map_.find(k);
        return it == map_.end() ? nullptr : it->second.get();
    }
    const toml::node* get(std::string_view k) const {
        auto it = map_.find(k);
        return it == map_.end() ? nullptr : it->second.get();
    }
    bool contains(std::string_view k) const { return map_.find(k) != map_.end(); }

    template <typename K, typename V> auto insert(K&& k, V&& v) {
        return map_.emplace(std::string(std::forward<K>(k)), impl::make_node(std::forward<V>(v)));
    }
    template <typename K, typename V> auto insert_or_assign(K&& k, V&& v) {
        auto key = std::string(std::forward<K>(k));
        map_[key] = impl::make_node(std::forward<V>(v));
        return map_.find(key);
    }
    template <typename K, typename V> auto emplace(K&& k, V&& v) {
        return insert(std::forward<K>(k), std::forward<V>(v));
    }
    void put_node(const std::string& k, std::unique_ptr<toml::node> p) { map_[k] = std::move(p); }
    size_t size() const { return map_.size(); }
    bool empty() const { return map_.empty(); }
    auto begin() const { return map_.begin(); }
    auto end() const { return map_.end(); }

  private:
    map_type map_;
};

inline const table* node::as_table() const { return dynamic_cast<const table*>(this); }
inline const array* node::as_array() const { return dynamic_cast<const array*>(this); }

namespace impl {
template <typename T> std::unique_ptr<node> make_node(T&& v) {
    using D = std::remove_cvref_t<T>;
    if constexpr (std::is_same_v<D, std::unique_ptr<node>>) {
        return v ? v->clone() : nullptr;
    } else if constexpr (std::is_base_of_v<node, D>) {
        return v.clone();
    } else if constexpr (std::is_same_v<D, bool>) {
        return std::make_unique<value<bool>>(v);
    } else if constexpr (std::is_integral_v<D>) {
        return std::make_unique<value<int64_t>>(static_cast<int64_t>(v));
    } else if constexpr (std::is_floating_point_v<D>) {
        return std::make_unique<value<double>>(static_cast<double>(v));
    } else if constexpr (std::is_same_v<D, date_time>) {
        return std::make_unique<value<date_time>>(v);
    } else {
        return std::make_unique<value<std::string>>(std::string(v));
    }
}
} // namespace impl

template <typename F> decltype(auto) node::visit(F&& f) const {
    if (auto* p = as<table>()) {
        return f(*p);
    }
    if (auto* p = as<array>()) {
        return f(*p);
    }
    if (auto* p = as<value<std::string>>()) {
        return f(*p);
    }
    if (auto* p = as<value<int64_t>>()) {
        return f(*p);
    }
    if (auto* p = as<value<double>>()) {
        return f(*p);
    }
    if (auto* p = as<value<bool>>()) {
        return f(*p);
    }
    return f(*as<value<date_time>>());
}

///// Printing (enough for round-tripping the reference's own save files).
inline void array::print(std::ostream& os, int indent) const {
    os << "[";
    bool first = true;
    for (const auto& e : elems_) {
        if (!first) {
            os << ", ";
        }
        first = false;
        if (e->is_table()) {
            os << "{ ";
            bool f2 = true;
            for (const auto& [k, v] : *e->as_table()) {
                if (!f2) {
                    os << ", ";
                }
                f2 = false;
                os << k << " = ";
                v->print(os, indent);
            }
            os << " }";
        } else {
            e->print(os, indent);
        }
    }
    os << "]";
}

namespace impl {
inline void print_table(std::ostream& os, const table& t, const std::string& prefix) {
    // Scalars and arrays first, then sub-tables.
    for (const auto& [k, v] : t) {
        if (!v->is_table()) {
            os << k << " = ";
            v->print(os, 0);
            os << "\n";
        }
    }
    for (const auto& [k, v] : t) {
        if (v->is_table()) {
            auto name = prefix.empty() ? k : prefix + "." + k;
            os << "\n[" << name << "]\n";
            print_table(os, *v->as_table(), name);
        }
    }
}
} // namespace impl
inline void table::print(std::ostream& os, int) const { impl::print_table(os, *this, ""); }

inline std::ostream& operator<<(std::ostream& os, const table& t) {
    t.print(os, 0);
    return os;
}
inline std::ostream& operator<<(std::ostream& os, const array& a) {
    a.print(os, 0);
    return os;
}
inline std::ostream& operator<<(std::ostream& os, const node& n) {
    n.print(os, 0);
    return os;
}

///// Parsing
class parse_error : public std::runtime_error {
  public:
    using std::runtime_error::runtime_error;
    std::string_view description() const { return what(); }
    source_region source() const { return {}; }
};

namespace impl {
class Parser {
  public:
    explicit Parser(std::string text)
        : s_(std::move(text)) {}

    table run() {
        table root;
        table* current = &root;
        while (true) {
            skip_ws_and_newlines();
            if (eof()) {
                break;
            }
            if (peek() == '[') {
                bool is_array = (pos_ + 1 < s_.size() && s_[pos_ + 1] == '[');
                pos_ += is_array ? 2 : 1;
                auto keys = parse_key_path();
                expect(']');
                if (is_array) {
                    expect(']');
                }
                current = open_table(root, keys, is_array);
            } else {
                auto keys = parse_key_path();
                skip_ws();
                expect('=');
                skip_ws();
                auto val = parse_value();
                table* t = current;
                for (size_t i = 0; i + 1 < keys.size(); ++i) {
                    t = descend(*t, keys[i]);
                }
                t->put_node(keys.back(), std::move(val));
            }
        }
        return root;
    }

  private:
    bool eof() const { return pos_ >= s_.size(); }
    char peek() const { return s_[pos_]; }
    [[noreturn]] void fail(const std::string& what) const {
        throw parse_error("toml shim: " + what + " at line " + std::to_string(line_));
    }
    void expect(char c) {
        skip_ws();
        if (eof() || s_[pos_] != c) {
            fail(std::string("expected '") + c + "'");
        }
        ++pos_;
    }
    void skip_ws() {
        while (!eof() && (peek() == ' ' || peek() == '\t')) {
            ++pos_;
        }
    }
    void skip_ws_and_newlines() {
        while (!eof()) {
            char c = peek();
            if (c == '#') {
                while (!eof() && peek() != '\n') {
                    ++pos_;
                }
            } else if (c == '\n') {
                ++line_;
                ++pos_;
            } else if (c == ' ' || c == '\t' || c == '\r') {
                ++pos_;
            } else {
                break;
            }
        }
    }
    std::string parse_quoted(char q) {
        ++pos_;
        std::string out;
        while (!eof() && peek() != q) {
            char c = s_[pos_++];
            if (c == '\\' && q == '"' && !eof()) {
                char e = s_[pos_++];
                switch (e) {
                    case 'n': out.push_back('\n'); break;
                    case 't': out.push_back('\t'); break;
                    case 'r': out.push_back('\r'); break;
                    default: out.push_back(e);
                }
            } else {
                out.push_back(c);
            }
        }
        if (eof()) {
            fail("unterminated string");
        }
        ++pos_;
        return out;
    }
    std::vector<std::string> parse_key_path() {
        std::vector<std::string> keys;
        while (true) {
            skip_ws();
            if (eof()) {
                fail("unexpected end of input in key");
            }
            if (peek() == '"' || peek() == '\'') {
                keys.push_back(parse_quoted(peek()));
            } else {
                size_t b = pos_;
                while (!eof() && (std::isalnum(static_cast<unsigned char>(peek())) ||
                                  peek() == '_' || peek() == '-')) {
                    ++pos_;
                }
                if (b == pos_) {
                    fail("empty key");
                }
                keys.emplace_back(s_.substr(b, pos_ - b));
            }
            skip_ws();
            if (!eof() && peek() == '.') {
                ++pos_;
                continue;
            }
            break;
        }
        return keys;
    }
    table* descend(table& t, const std::string& key) {
        auto* n = t.get(key);
        if (n == nullptr) {
            auto fresh = std::make_unique<table>();
            fresh->set_line(line_);
            t.put_node(key, std::move(fresh));
            n = t.get(key);
        }
        if (auto* sub = n->as<table>()) {
            return sub;
        }
        if (auto* arr = n->as<array>()) {
            if (arr->empty()) {
                fail("empty array used as a table path");
            }
            if (auto* last = arr->back().as<table>()) {
                return last;
            }
        }
        fail("key '" + key + "' is not a table");
    }
    table* open_table(table& root, const std::vector<std::string>& keys, bool is_array) {
        table* t = &root;
        for (size_t i = 0; i + 1 < keys.size(); ++i) {
            t = descend(*t, keys[i]);
        }
        const auto& last = keys.back();
        if (!is_array) {
            return descend(*t, last);
        }
        auto* n = t->get(last);
        if (n == nullptr) {
            auto fresh = std::make_unique<array>();
            fresh->set_line(line_);
            t->put_node(last, std::move(fresh));
            n = t->get(last);
        }
        auto* arr = n->as<array>();
        if (arr == nullptr) {
            fail("key '" + last + "' is not an array of tables");
        }
        auto fresh = std::make_unique<table>();
        fresh->set_line(line_);
        arr->emplace_back_node(std::move(fresh));
        return arr->back().as<table>();
    }
    std::unique_ptr<node> parse_value() {
        skip_ws();
        if (eof()) {
            fail("missing value");
        }
        std::unique_ptr<node> out;
        char c = peek();
        if (c == '"' || c == '\'') {
            out = std::make_unique<value<std::string>>(parse_quoted(c));
        } else if (c == '[') {
            ++pos_;
            auto arr = std::make_unique<array>();
            while (true) {
                skip_ws_and_newlines();
                if (eof()) {
                    fail("unterminated array");
                }
                if (peek() == ']') {
                    ++pos_;
                    break;
                }
                arr->emplace_back_node(parse_value());
                skip_ws_and_newlines();
                if (!eof() && peek() == ',') {
                    ++pos_;
                }
            }
            out = std::move(arr);
        } else if (c == '{') {
            ++pos_;
            auto tab = std::make_unique<table>();
            while (true) {
                skip_ws();
                if (eof()) {
                    fail("unterminated inline table");
                }
                if (peek() == '}') {
                    ++pos_;
                    break;
                }
                auto keys = parse_key_path();
                expect('=');
                auto v = parse_value();
                table* t = tab.get();
                for (size_t i = 0; i + 1 < keys.size(); ++i) {
                    t = descend(*t, keys[i]);
                }
                t->put_node(keys.back(), std::move(v));
                skip_ws();
                if (!eof() && peek() == ',') {
                    ++pos_;
                }
            }
            out = std::move(tab);
        } else {
            size_t b = pos_;
            while (!eof() && peek() != ',' && peek() != ']' && peek() != '}' &&
                   peek() != '\n' && peek() != '#' && peek() != '\r') {
                ++pos_;
            }
            std::string tok = s_.substr(b, pos_ - b);
            while (!tok.empty() && (tok.back() == ' ' || tok.back() == '\t')) {
                tok.pop_back();
            }
            out = parse_scalar(tok);
        }
        out->set_line(line_);
        return out;
    }
    std::unique_ptr<node> parse_scalar(const std::string& tok) {
        if (tok == "true") {
            return std::make_unique<value<bool>>(true);
        }
        if (tok == "false") {
            return std::make_unique<value<bool>>(false);
        }
        if (tok.empty()) {
            fail("empty value");
        }
        // Local date-time: YYYY-MM-DDTHH:MM:SS
        if (tok.size() >= 19 && tok[4] == '-' && tok[7] == '-' &&
            (tok[10] == 'T' || tok[10] == ' ') && tok[13] == ':') {
            date d(
                std::atoi(tok.substr(0, 4).c_str()),
                unsigned(std::atoi(tok.substr(5, 2).c_str())),
                unsigned(std::atoi(tok.substr(8, 2).c_str()))
            );
            time t{
                uint8_t(std::atoi(tok.substr(11, 2).c_str())),
                uint8_t(std::atoi(tok.substr(14, 2).c_str())),
                uint8_t(std::atoi(tok.substr(17, 2).c_str())),
                0};
            return std::make_unique<value<date_time>>(date_time(d, t));
        }
        std::string clean;
        for (char ch : tok) {
            if (ch != '_') {
                clean.push_back(ch);
            }
        }
        bool is_float = clean.find_first_of(".eE") != std::string::npos ||
                        clean.find("inf") != std::string::npos ||
                        clean.find("nan") != std::string::npos;
        bool is_hex = clean.rfind("0x", 0) == 0;
        char* endp = nullptr;
        if (is_float && !is_hex) {
            double v = std::strtod(clean.c_str(), &endp);
            if (endp == clean.c_str() || *endp != '\0') {
                fail("bad float '" + tok + "'");
            }
            return std::make_unique<value<double>>(v);
        }
        long long v = std::strtoll(clean.c_str(), &endp, is_hex ? 16 : 10);
        if (endp == clean.c_str() || *endp != '\0') {
            fail("bad value '" + tok + "'");
        }
        return std::make_unique<value<int64_t>>(int64_t(v));
    }

    std::string s_;
    size_t pos_ = 0;
    uint32_t line_ = 1;
};
} // namespace impl

inline table parse(std::string_view text, std::string_view = {}) {
    return impl::Parser(std::string(text)).run();
}
inline table parse(std::istream& is, std::string_view = {}) {
    std::stringstream ss;
    ss << is.rdbuf();
    return impl::Parser(ss.str()).run();
}
inline table parse_file(std::string_view path) {
    std::ifstream f{std::string(path)};
    if (!f) {
        throw parse_error("toml shim: cannot open " + std::string(path));
    }
    return parse(f);
}
} // namespace toml
