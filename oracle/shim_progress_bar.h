// TEST INFRASTRUCTURE ONLY (oracle build).  Force-included ahead of the reference sources with
// -DPROGRESS_BAR so that /root/reference/src/progress_bar.h (which does not compile on GCC 13.3:
// it hands rvalues to std::make_format_args at progress_bar.h:40) is replaced by a silent class of
// the same shape.  No reference file is edited; the progress bar is cosmetic.
#pragma once
// progress_bar.h is also what pulls <format>/<cstdio> into src/index.cpp (progress_bar.h:10-11)
#include <cstdio>
#include <format>
#include <string>
class progress_bar {
public:
    progress_bar() = default;
    explicit progress_bar(const std::string&) {}
    void update(double) {}
};
