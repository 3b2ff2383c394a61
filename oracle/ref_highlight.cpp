// TEST INFRASTRUCTURE ONLY.  Exposes the reference highlighter through a C ABI.
//
// `ac_automaton` is a file-local class of /root/reference/src/database.cpp (lines 26-138), so this
// translation unit *includes that source where it lies* (nothing is copied) and calls
// ac_automaton::render (database.cpp:58-91) directly.  Built into oracle/_ref/libcoffeeref.so.
#include "database.cpp"  // resolved through -I/root/reference/src

#include <cstdlib>
#include <cstring>

extern "C" {

// Renders `text` with every occurrence of the keyword set wrapped in left/right, exactly as
// select(...highlight...) does for one string field (database.cpp:139-165, 394-432).
// Returns the length of the rendered string; *out is malloc'd (free with ref_free).
int64_t ref_render(const char* kw_bytes, const int64_t* kw_off, int64_t nkw, const char* text, int64_t tlen,
                   const char* left, int64_t llen, const char* right, int64_t rlen, char** out) {
    std::vector<std::string> keywords;
    for (int64_t k = 0; k < nkw; ++k) {
        keywords.emplace_back(kw_bytes + kw_off[k], (size_t)(kw_off[k + 1] - kw_off[k]));
    }
    ac_automaton ac(keywords);
    std::string res = ac.render(std::string(text, (size_t)tlen), std::string(left, (size_t)llen),
                                std::string(right, (size_t)rlen));
    char* buf = (char*)std::malloc(res.size() + 1);
    std::memcpy(buf, res.data(), res.size());
    buf[res.size()] = 0;
    *out = buf;
    return (int64_t)res.size();
}

}  // extern "C"
