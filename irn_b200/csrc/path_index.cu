// R1/R2: PathIndex tables on the host (integer, bit-exact with misc/indexing.py:6-88) and the
// compact path table the device kernels consume.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "common.h"
#include "path_tables.h"

namespace irn {

static thread_local char g_err[512];
static thread_local int g_launches = 0;
char* last_error_buf() { return g_err; }
int& launch_counter() { return g_launches; }
static thread_local long long g_total_launches = 0;
long long& total_launch_counter() { return g_total_launches; }

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

// Enumerate the half-plane destinations and their straight-line paths.
//   misc/indexing.py:22-31  destinations: (0,x) x=1..r-1, then y=1..r-1, |x|<r, x^2+y^2<r^2
//   misc/indexing.py:33-48  path = bounding-box points with cross^2 < len^2, stably sorted by
//                           descending L1 norm (destination first, source last)
//   misc/indexing.py:50-54  grouped by path length, ascending
PathTable build_path_table(int radius) {
    PathTable t;
    t.radius = radius;
    std::vector<std::pair<int, int>> dirs;
    for (int x = 1; x < radius; ++x) dirs.push_back({0, x});
    for (int y = 1; y < radius; ++y)
        for (int x = -radius + 1; x < radius; ++x)
            if (x * x + y * y < radius * radius) dirs.push_back({y, x});

    std::vector<std::vector<std::pair<int, int>>> paths(dirs.size());
    int max_len = 0;
    for (size_t i = 0; i < dirs.size(); ++i) {
        const int dy = dirs[i].first, dx = dirs[i].second;
        const long len_sq = (long)dy * dy + (long)dx * dx;
        std::vector<std::pair<int, int>>& p = paths[i];
        for (int y = std::min(0, dy); y <= std::max(0, dy); ++y)
            for (int x = std::min(0, dx); x <= std::max(0, dx); ++x) {
                const long cross = (long)dy * x - (long)dx * y;
                if (cross * cross < len_sq) p.push_back({y, x});
            }
        std::stable_sort(p.begin(), p.end(), [](const std::pair<int, int>& a, const std::pair<int, int>& b) {
            return std::abs(a.first) + std::abs(a.second) > std::abs(b.first) + std::abs(b.second);
        });
        max_len = std::max(max_len, (int)p.size());
    }
    for (int L = 1; L <= max_len; ++L) {
        int n = 0;
        for (size_t i = 0; i < dirs.size(); ++i)
            if ((int)paths[i].size() == L) {
                t.dst.push_back(dirs[i]);
                t.path_start.push_back((int)t.points.size());
                for (auto& q : paths[i]) t.points.push_back(q);
                ++n;
            }
        if (n) {
            t.group_len.push_back(L);
            t.group_paths.push_back(n);
        }
    }
    t.path_start.push_back((int)t.points.size());
    return t;
}

}  // namespace irn

using namespace irn;

extern "C" const char* irn_last_error(void) { return last_error_buf(); }
extern "C" int irn_version(void) { return 100; }
extern "C" long long irn_total_launch_count(void) { return total_launch_counter(); }

extern "C" int irn_path_index_shape(int radius, int* n_dst, int* n_groups, int* group_len, int* group_paths) {
    if (radius < 2 || radius > 64) return fail(kBadArg, "irn_path_index_shape: radius %d out of range [2,64]", radius);
    PathTable t = build_path_table(radius);
    if (n_dst) *n_dst = (int)t.dst.size();
    if (n_groups) *n_groups = (int)t.group_len.size();
    for (size_t g = 0; g < t.group_len.size(); ++g) {
        if (group_len) group_len[g] = t.group_len[g];
        if (group_paths) group_paths[g] = t.group_paths[g];
    }
    return kOk;
}

extern "C" int irn_path_index_fill(int radius, int Hp, int Wp, int64_t* search_dst, int64_t* search_paths,
                                   int64_t* path_indices, int64_t* src_indices, int64_t* dst_indices) {
    if (radius < 2 || radius > 64) return fail(kBadArg, "irn_path_index_fill: radius %d out of range [2,64]", radius);
    const int rf = radius - 1;  // ceil(r) - 1 for integer r (misc/indexing.py:10)
    const int ch = Hp - rf, cw = Wp - 2 * rf;
    if (ch <= 0 || cw <= 0) return fail(kBadArg, "irn_path_index_fill: grid %dx%d too small for radius %d", Hp, Wp, radius);
    PathTable t = build_path_table(radius);
    const int64_t n_src = (int64_t)ch * cw;
    const int n_dst = (int)t.dst.size();

    if (search_dst)
        for (int k = 0; k < n_dst; ++k) {
            search_dst[2 * k] = t.dst[k].first;
            search_dst[2 * k + 1] = t.dst[k].second;
        }
    if (search_paths) {
        int64_t* o = search_paths;
        for (auto& q : t.points) {
            *o++ = q.first;
            *o++ = q.second;
        }
    }
    // misc/indexing.py:82: src = full[:ch, rf:rf+cw]
    if (src_indices)
        for (int y = 0; y < ch; ++y)
            for (int x = 0; x < cw; ++x) src_indices[(int64_t)y * cw + x] = (int64_t)y * Wp + rf + x;
    // misc/indexing.py:66-80: per path point, full[dy:dy+ch, rf+dx:rf+dx+cw] flattened
    int64_t* out = path_indices;
    for (int k = 0; k < n_dst; ++k) {
        for (int j = t.path_start[k]; j < t.path_start[k + 1]; ++j) {
            const int64_t off = (int64_t)t.points[j].first * Wp + t.points[j].second;
            if (j == t.path_start[k] && dst_indices)
                for (int y = 0; y < ch; ++y)
                    for (int x = 0; x < cw; ++x)
                        dst_indices[(int64_t)k * n_src + (int64_t)y * cw + x] = (int64_t)y * Wp + rf + x + off;
            if (path_indices) {
                for (int y = 0; y < ch; ++y)
                    for (int x = 0; x < cw; ++x) out[(int64_t)y * cw + x] = (int64_t)y * Wp + rf + x + off;
                out += n_src;
            }
        }
    }
    return kOk;
}
