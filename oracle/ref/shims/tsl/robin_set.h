// Build shim (test infrastructure, see oracle/README.md): the reference includes
// tessil robin-map 1.4.0 via CMake FetchContent, which is unreachable offline.
// The search/build paths only need an unordered set, so alias the STL one.
#pragma once
#include <functional>
#include <memory>
#include <unordered_set>
namespace tsl {
template <
    class Key,
    class Hash = std::hash<Key>,
    class Eq = std::equal_to<Key>,
    class Alloc = std::allocator<Key>>
using robin_set = std::unordered_set<Key, Hash, Eq, Alloc>;
} // namespace tsl
