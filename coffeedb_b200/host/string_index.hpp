// coffeedb_b200::string_index — C++ adaptor that puts the B200 engine (C ABI, include/coffeedb_b200.h) behind
// the reference's `string_index` surface (src/index.h:54-86 of sunkafei/coffeedb):
//
//     void add(int64_t id, std::string_view value);                                   // src/index.cpp:174-177
//     void build() override;                                                          // src/index.cpp:178-236
//     std::vector<std::pair<int64_t,int64_t>> query(const std::string&) const override;  // src/index.cpp:237-326
//
// Same argument meaning, same result layout (pairs of (id, count) in ascending doc index), same exceptions
// (std::runtime_error with the reference's three messages; CUDA failures surface the same way, so the server's
// catch at src/server.cpp:58-62 still answers HTTP 500).  Non-copyable / non-movable like `index`
// (src/index.h:12-15).  There is no CPU fallback.
//
// The class is a template over its base so that it can derive from the reference's own abstract `index`
// (`using string_index = coffeedb_b200::basic_string_index<index>;` inside src/index.h) and
// still be compiled and tested stand-alone, where `coffeedb_b200::index_base` supplies the same virtuals.
//
// Additions that the reference does not have (all optional):
//   query_batch(keywords)  one device pass for the keyword list of one key (src/interface.cpp:79-86 loops
//                          query() per keyword) — rows in keyword order
//   spans(keywords, docs)  merged highlight spans from suffix-array positions (replaces the span loop of
//                          ac_automaton::render, src/database.cpp:58-77); render() splices the markers
//                          (src/database.cpp:78-90)
#pragma once
#include <algorithm>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <string_view>
#include <utility>
#include <vector>

#include "../../include/coffeedb_b200.h"

namespace coffeedb_b200 {

// Stand-alone base with the virtuals of the reference's `index` (src/index.h:9-23).
class index_base {
public:
    index_base() {}
    index_base(const index_base&) = delete;
    index_base(index_base&&) = delete;
    index_base& operator=(const index_base&) = delete;
    index_base& operator=(index_base&&) = delete;
    virtual std::vector<std::pair<int64_t, int64_t>> query(const std::string&) const {
        throw std::logic_error("Unimplemented method index::query");
    }
    virtual void build() { throw std::logic_error("Unimplemented method index::build"); }
    virtual ~index_base() {}
};

template <class Base>
class basic_string_index : public Base {
public:
    using value_type = std::string;
    static constexpr int8_t number = 3;  // src/index.h:76: on-disk type tag of string fields
    using result_type = std::vector<std::pair<int64_t, int64_t>>;

    explicit basic_string_index(int device = -1, bool compat_signed = true) {
        cdb_options o{};
        o.device = device;
        o.compat_signed = compat_signed ? 1 : 0;
        check(cdb_create(&o, &h_));
    }
    ~basic_string_index() override { cdb_destroy(h_); }

    void add(int64_t id, std::string_view value) { check(cdb_add(h_, id, value.data(), (int64_t)value.size())); }

    void build() override { check(cdb_build(h_)); }

    // build() that reads the suffix array back from `path` when the file was saved for this very corpus (SURVEY.md 8f-4;
    // the reference rebuilds at every start, src/server.cpp:43-44).  Returns true when the array came from the file.
    bool build_or_load(const std::string& path) {
        int32_t loaded = 0;
        check(cdb_build_or_load(h_, path.c_str(), &loaded));
        return loaded != 0;
    }
    void save(const std::string& path) const { check(cdb_save(h_, path.c_str())); }

    // One keyword, as the server calls it (src/database.cpp:387-393).  Concurrent calls on the same index — the
    // reference runs query() from up to max(8, hw-1) httplib threads — are coalesced into one device batch by
    // cdb_query (micro_batcher.hpp); a lone call is served at once.
    result_type query(const std::string& keyword) const override {
        cdb_result r{};
        check(cdb_query(h_, keyword.data(), (int64_t)keyword.size(), &r));
        result_type out = row(r, 0);
        cdb_result_free(&r);
        return out;
    }

    // One batched device pass for several keywords; row q is what query(keywords[q]) returns.
    std::vector<result_type> query_batch(const std::vector<std::string>& keywords) const {
        std::string bytes;
        std::vector<int64_t> off(keywords.size() + 1, 0);
        for (size_t q = 0; q < keywords.size(); ++q) {
            bytes += keywords[q];
            off[q + 1] = (int64_t)bytes.size();
        }
        cdb_result r{};
        check(cdb_locate_batch(h_, bytes.data(), off.data(), (int64_t)keywords.size(), &r));
        std::vector<result_type> out(keywords.size());
        for (size_t q = 0; q < keywords.size(); ++q) out[q] = row(r, (int64_t)q);
        cdb_result_free(&r);
        return out;
    }

    // Inclusive [begin, end] spans of `keywords` in each of the documents `docs` (doc indices = add() order).
    std::vector<result_type> spans(const std::vector<std::string>& keywords, const std::vector<int64_t>& docs) const {
        std::string bytes;
        std::vector<int64_t> off(keywords.size() + 1, 0);
        for (size_t q = 0; q < keywords.size(); ++q) {
            bytes += keywords[q];
            off[q + 1] = (int64_t)bytes.size();
        }
        cdb_spans s{};
        check(cdb_locate_spans(h_, bytes.data(), off.data(), (int64_t)keywords.size(), docs.data(), (int64_t)docs.size(), &s));
        std::vector<result_type> out(docs.size());
        for (size_t t = 0; t < docs.size(); ++t)
            for (int64_t i = s.span_off[t]; i < s.span_off[t + 1]; ++i) out[t].emplace_back(s.spans[2 * i], s.spans[2 * i + 1]);
        cdb_spans_free(&s);
        return out;
    }

    // Marker splicing of src/database.cpp:78-90 over spans computed by spans().
    static std::string render(std::string_view text, const result_type& sp, std::string_view left, std::string_view right) {
        std::vector<int64_t> flat;
        flat.reserve(sp.size() * 2);
        for (auto& p : sp) {
            flat.push_back(p.first);
            flat.push_back(p.second);
        }
        std::string out((size_t)cdb_splice(text.data(), (int64_t)text.size(), flat.data(), (int64_t)sp.size(), left.data(),
                                           (int64_t)left.size(), right.data(), (int64_t)right.size(), nullptr, 0), '\0');
        cdb_splice(text.data(), (int64_t)text.size(), flat.data(), (int64_t)sp.size(), left.data(), (int64_t)left.size(),
                   right.data(), (int64_t)right.size(), out.data(), (int64_t)out.size());
        return out;
    }

    cdb_index* handle() const { return h_; }

private:
    static void check(cdb_status s) {
        if (s != CDB_OK) throw std::runtime_error(cdb_last_error());
    }
    static result_type row(const cdb_result& r, int64_t q) {
        result_type out;
        out.reserve((size_t)(r.row_off[q + 1] - r.row_off[q]));
        for (int64_t i = r.row_off[q]; i < r.row_off[q + 1]; ++i) out.emplace_back(r.pairs[2 * i], r.pairs[2 * i + 1]);
        return out;
    }
    cdb_index* h_ = nullptr;
};

using string_index = basic_string_index<index_base>;

// The same surface over several GPUs of this process (cdb_sharded_*, include/coffeedb_b200.h): documents are split into
// contiguous doc-index ranges, one shard per entry of `devices`; query() rows are the whole-corpus rows (ascending doc
// index).  `using string_index = coffeedb_b200::basic_sharded_string_index<index>;` in src/index.h, constructed with the
// device list, shards every string key of the database over the box's GPUs without touching database.cpp.
template <class Base>
class basic_sharded_string_index : public Base {
public:
    using value_type = std::string;
    static constexpr int8_t number = 3;
    using result_type = std::vector<std::pair<int64_t, int64_t>>;

    explicit basic_sharded_string_index(const std::vector<int32_t>& devices = default_devices(), bool compat_signed = true) {
        cdb_options o{};
        o.device = -1;
        o.compat_signed = compat_signed ? 1 : 0;
        check(cdb_sharded_create(devices.data(), (int32_t)devices.size(), &o, &h_));
    }
    ~basic_sharded_string_index() override { cdb_sharded_destroy(h_); }

    void add(int64_t id, std::string_view value) { check(cdb_sharded_add(h_, id, value.data(), (int64_t)value.size())); }
    void build() override { check(cdb_sharded_build(h_)); }

    result_type query(const std::string& keyword) const override { return query_batch({keyword})[0]; }

    std::vector<result_type> query_batch(const std::vector<std::string>& keywords) const {
        std::string bytes;
        std::vector<int64_t> off(keywords.size() + 1, 0);
        for (size_t q = 0; q < keywords.size(); ++q) {
            bytes += keywords[q];
            off[q + 1] = (int64_t)bytes.size();
        }
        cdb_result r{};
        check(cdb_sharded_locate_batch(h_, bytes.data(), off.data(), (int64_t)keywords.size(), &r));
        std::vector<result_type> out(keywords.size());
        for (size_t q = 0; q < keywords.size(); ++q) {
            out[q].reserve((size_t)(r.row_off[q + 1] - r.row_off[q]));
            for (int64_t i = r.row_off[q]; i < r.row_off[q + 1]; ++i) out[q].emplace_back(r.pairs[2 * i], r.pairs[2 * i + 1]);
        }
        cdb_result_free(&r);
        return out;
    }

    // every CUDA device of the box, one shard each
    static std::vector<int32_t> default_devices() {
        std::vector<int32_t> d;
        for (int g = 0; g < cdb_device_count(); ++g) d.push_back(g);
        if (d.empty()) d.push_back(0);
        return d;
    }

private:
    static void check(cdb_status s) {
        if (s != CDB_OK) throw std::runtime_error(cdb_last_error());
    }
    cdb_sharded* h_ = nullptr;
};

using sharded_string_index = basic_sharded_string_index<index_base>;

// $correlation composition over (id, count) lists, as filter() does it on the host (src/interface.cpp:79-146).  The
// device delivers exact per-(keyword, document) counts; these helpers are the reference-shaped integer adds on top.
namespace correlation {
using list = std::vector<std::pair<int64_t, int64_t>>;

// OR of two keyword lists of one key: union by id, counts added (src/interface.cpp:87-113).  Inputs sorted by (id, count).
inline list or_merge(const list& now, const list& result) {
    list tmp;
    size_t i = 0, j = 0;
    while (i < now.size() && j < result.size()) {
        if (now[i].first == result[j].first) {
            tmp.push_back(now[i]);
            tmp.back().second += result[j].second;
            ++i, ++j;
        } else if (now[i] < result[j]) {
            tmp.push_back(now[i++]);
        } else {
            tmp.push_back(result[j++]);
        }
    }
    while (i < now.size()) tmp.push_back(now[i++]);
    while (j < result.size()) tmp.push_back(result[j++]);
    return tmp;
}

// AND across keys: intersection by id, counts added (src/interface.cpp:118-135).
inline list and_merge(const list& result, const list& answer) {
    list tmp;
    for (size_t i = 0, j = 0; i < result.size() && j < answer.size();) {
        if (result[i].first == answer[j].first) {
            tmp.push_back(result[i]);
            tmp.back().second += answer[j].second;
            ++i, ++j;
        } else if (result[i] < answer[j]) {
            ++i;
        } else {
            ++j;
        }
    }
    return tmp;
}

// The keyword list of ONE key (src/interface.cpp:79-113) from the rows of one batched device pass: every row is
// sorted by (id, count) first, as filter() does on arrival (:81-82, :86-87), then OR-merged in keyword order.
inline list or_of_rows(std::vector<list> rows) {
    list result;
    for (size_t k = 0; k < rows.size(); ++k) {
        std::sort(rows[k].begin(), rows[k].end());
        result = k == 0 ? std::move(rows[k]) : or_merge(rows[k], result);
    }
    return result;
}

// $correlation range [L, R) and the final descending sort (src/interface.cpp:137-146; std::sort with the same
// comparator on the same input order gives the same permutation for a given libstdc++).
inline void finish(list& answer, bool has_range, int64_t L, int64_t R) {
    if (has_range)
        answer.erase(std::remove_if(answer.begin(), answer.end(), [L, R](auto p) { return !(p.second >= L && p.second < R); }),
                     answer.end());
    std::sort(answer.begin(), answer.end(), [](auto x, auto y) { return x.second > y.second; });
}
}  // namespace correlation

}  // namespace coffeedb_b200
