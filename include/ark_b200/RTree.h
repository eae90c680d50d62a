// ark_b200/RTree.h -- the data side of ark::RTree (include/RTree.h:15-176 of the reference) for the C++ facade: the
// node array, the leaf distributions, leafBestMatch and the part map, read from the reference's model files
// (RTree::loadFile RTree.cpp:2967-3061: binary 'R' format and legacy text; readPartMap :3465-3510;
// updateBestMatchTable :3452-3463).  Prediction runs on the device: attach() hands the tree to a fitter
// (avb_fitter_set_rtree), after which avb_rtree_predict_batch / avb_upload_depth_batch(parts = NULL) use it.
// Training, postProcess and the cv::Mat front ends are not part of the facade.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

struct avb_fitter;

namespace ark {

class RTree {
public:
    struct RNode {                 // include/RTree.h:28-41
        float u[2] = {0, 0}, v[2] = {0, 0};
        float thresh = 0;
        int lnode = -1, rnode = -1;
        int leafid = -1;
    };
    explicit RTree(int num_parts) : numParts(num_parts) {}
    explicit RTree(const std::string& path) { loadFile(path); }
    bool loadFile(const std::string& path);          // false: unreadable file; malformed files throw std::runtime_error
    void updateBestMatchTable();
    static bool readPartMap(const std::string& path, std::vector<int>& result, int& num_new_parts, int& partmap_type);
    void attach(avb_fitter* fitter) const;           // throws std::runtime_error on failure

    std::vector<RNode> nodes;
    std::vector<std::vector<float>> leafData;        // [leaf][numParts]
    std::vector<uint8_t> leafBestMatch;
    int numParts = 0;
    std::vector<int> partMap;
    int partMapType = -1;
};

}  // namespace ark
