// Build shim (test infrastructure, see oracle/README.md): stand-in for tessil
// robin-map 1.4.0. The only robin-specific API the reference touches on the Vamana
// search/build paths is `iterator.key()` / `iterator.value()`, added here on top of
// std::unordered_map.
#pragma once
#include <functional>
#include <memory>
#include <unordered_map>
#include <utility>
namespace tsl {
template <
    class Key,
    class T,
    class Hash = std::hash<Key>,
    class Eq = std::equal_to<Key>,
    class Alloc = std::allocator<std::pair<const Key, T>>>
class robin_map : public std::unordered_map<Key, T, Hash, Eq, Alloc> {
    using base = std::unordered_map<Key, T, Hash, Eq, Alloc>;

  public:
    using base::base;
    using const_iterator = typename base::const_iterator;

    struct iterator : base::iterator {
        iterator() = default;
        iterator(typename base::iterator it)
            : base::iterator(it) {}
        const Key& key() const { return (**this).first; }
        T& value() const { return (**this).second; }
    };

    template <class... Args>
    std::pair<iterator, bool> try_emplace(const Key& k, Args&&... args) {
        auto r = base::try_emplace(k, std::forward<Args>(args)...);
        return {iterator(r.first), r.second};
    }
    iterator find(const Key& k) { return iterator(base::find(k)); }
    const_iterator find(const Key& k) const { return base::find(k); }
    iterator begin() { return iterator(base::begin()); }
    iterator end() { return iterator(base::end()); }
    const_iterator begin() const { return base::begin(); }
    const_iterator end() const { return base::end(); }
};
} // namespace tsl
