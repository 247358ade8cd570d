// Compact PathIndex description shared by the host API (path_index.cu) and the device
// constant tables (rw.cu).
#pragma once
#include <utility>
#include <vector>

namespace irn {

struct PathTable {
    int radius = 0;
    std::vector<std::pair<int, int>> dst;     // (dy,dx) per destination, reference order (grouped by path length)
    std::vector<int> path_start;              // n_dst+1 offsets into `points`
    std::vector<std::pair<int, int>> points;  // path points, destination first, source (0,0) last
    std::vector<int> group_len, group_paths;  // per length group
};

PathTable build_path_table(int radius);

}  // namespace irn
