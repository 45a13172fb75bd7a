// oracle/shim/boost/multi_index_container.hpp -- TEST INFRASTRUCTURE, not product code.
//
// Stand-in for the exact Boost.MultiIndex subset that the reference's PairCount.h uses
// (PairCount.h:84-92 and :211-219): a container with
//     index 0 = hashed_unique<member<V, pair<T,T>, &V::pair>>
//     index 1 = ordered_non_unique<identity<V>, Cmp>
// Boost is not installed in this image, so the reference headers are compiled verbatim
// against this std-container implementation (std::list nodes + unordered_map + multiset
// of node pointers). Only observable semantics matter: unique key lookup, modify(),
// and begin() of the ordered index. Written from SURVEY.md section 8(c); not Boost code.
#pragma once
#include <cstddef>
#include <functional>
#include <list>
#include <set>
#include <unordered_map>
#include <utility>

namespace boost {
namespace multi_index {

template <class C, class T, T C::*P>
struct member {
    using result_type = T;
    const T &operator()(const C &c) const { return c.*P; }
};
template <class V>
struct identity {
    using result_type = V;
    const V &operator()(const V &v) const { return v; }
};
template <class KE>
struct hashed_unique {
    using key_extractor = KE;
};
template <class KE, class Cmp>
struct ordered_non_unique {
    using key_extractor = KE;
    using compare = Cmp;
};
template <class... I>
struct indexed_by {};

namespace shim_detail {
template <class K>
struct key_hash {
    size_t operator()(const K &k) const { return std::hash<K>()(k); }
};
template <class A, class B>
struct key_hash<std::pair<A, B>> {
    size_t operator()(const std::pair<A, B> &p) const {
        size_t seed = 0;
        seed ^= std::hash<A>()(p.first) + 0x9e3779b9 + (seed << 6) + (seed >> 2);
        seed ^= std::hash<B>()(p.second) + 0x9e3779b9 + (seed << 6) + (seed >> 2);
        return seed;
    }
};
}  // namespace shim_detail

template <class V, class IB>
class multi_index_container;

template <class V, class KE0, class KE1, class Cmp>
class multi_index_container<V, indexed_by<hashed_unique<KE0>, ordered_non_unique<KE1, Cmp>>> {
    struct ByValue {
        bool operator()(const V *a, const V *b) const { return Cmp()(*a, *b); }
    };
    using Ordered = std::multiset<const V *, ByValue>;
    struct Node {
        V value;
        typename Ordered::iterator ord;
        explicit Node(const V &v) : value(v) {}
    };
    using Nodes = std::list<Node>;
    using Key = typename KE0::result_type;

    Nodes nodes_;
    std::unordered_map<Key, typename Nodes::iterator, shim_detail::key_hash<Key>> by_key_;
    Ordered ordered_;

  public:
    class iterator {
        typename Nodes::const_iterator it_;
        friend class multi_index_container;

      public:
        iterator() = default;
        explicit iterator(typename Nodes::const_iterator i) : it_(i) {}
        const V &operator*() const { return it_->value; }
        const V *operator->() const { return &it_->value; }
        iterator &operator++() {
            ++it_;
            return *this;
        }
        bool operator==(const iterator &o) const { return it_ == o.it_; }
        bool operator!=(const iterator &o) const { return it_ != o.it_; }
    };
    class ordered_iterator {
        typename Ordered::const_iterator it_;

      public:
        explicit ordered_iterator(typename Ordered::const_iterator i) : it_(i) {}
        const V &operator*() const { return **it_; }
        ordered_iterator &operator++() {
            ++it_;
            return *this;
        }
        bool operator!=(const ordered_iterator &o) const { return it_ != o.it_; }
    };

    struct hashed_view {
        multi_index_container *c;
        iterator find(const Key &k) const {
            auto f = c->by_key_.find(k);
            return f == c->by_key_.end() ? c->end() : iterator(f->second);
        }
        template <class F>
        bool modify(iterator pos, F f) {
            auto li = c->nodes_.erase(pos.it_, pos.it_);  // const_iterator -> iterator, erases nothing
            auto handle = c->ordered_.extract(li->ord);   // relink without free+malloc
            f(li->value);
            li->ord = c->ordered_.insert(std::move(handle));
            return true;
        }
    };
    struct ordered_view {
        const multi_index_container *c;
        bool empty() const { return c->ordered_.empty(); }
        ordered_iterator begin() const { return ordered_iterator(c->ordered_.begin()); }
        ordered_iterator end() const { return ordered_iterator(c->ordered_.end()); }
    };

    multi_index_container() = default;
    multi_index_container(const multi_index_container &) = delete;
    multi_index_container &operator=(const multi_index_container &) = delete;

    template <int N>
    auto &get() {
        if constexpr (N == 0)
            return hashed_;
        else
            return ordered_view_;
    }
    template <int N>
    const auto &get() const {
        static_assert(N == 1, "only the ordered index is read through a const container");
        return ordered_view_;
    }

    size_t size() const { return nodes_.size(); }
    iterator begin() const { return iterator(nodes_.begin()); }
    iterator end() const { return iterator(nodes_.end()); }

    std::pair<iterator, bool> insert(const V &v) {
        Key k = KE0()(v);
        auto f = by_key_.find(k);
        if (f != by_key_.end()) return {iterator(f->second), false};
        nodes_.emplace_back(v);
        auto li = std::prev(nodes_.end());
        li->ord = ordered_.insert(&li->value);
        by_key_.emplace(k, li);
        return {iterator(li), true};
    }

  private:
    hashed_view hashed_{this};
    ordered_view ordered_view_{this};
};

}  // namespace multi_index
using multi_index::multi_index_container;
}  // namespace boost
