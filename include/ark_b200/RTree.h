// ark_b200/RTree.h -- the data side of ark::RTree (include/RTree.h:15-176 of the reference) for the C++ facade: the
// node array, the leaf distributions, leafBestMatch and the part map, read from the reference's model files
// (RTree::loadFile RTree.cpp:2967-3061: binary 'R' format and legacy text; readPartMap :3465-3510;
// updateBestMatchTable :3452-3463).  Prediction runs on the device: attach() hands the tree to a fitter
// (avb_fitter_set_rtree), after which avb_rtree_predict_batch / avb_upload_depth_batch(parts = NULL) use it.
// predictBest / postProcess are the reference's cv::Mat front ends over avb_rtree_predict_batch /
// avb_rtree_postprocess_batch (the tree owns a small private fitter for them).  Training is not part of the facade.
#pragma once
#include <cstdint>
#include <string>
#include <vector>
#include "AvatarOptimizer.h"   // Eigen / OpenCV types (or their stand-ins)

struct avb_fitter;
struct avb_model;

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
    ~RTree();
    /** RTree.cpp:3184-3262: CV_8U labels (255 = not predicted) of a CV_32F depth image, on the device */
    cv::Mat predictBest(const cv::Mat& depth, int num_threads = 1, int interval = 1, cv::Point top_left = cv::Point(0, 0),
                        cv::Point bot_right = cv::Point(-1, -1), bool fill_in_gaps = true);
    /** RTree.cpp:3422-3450, on the device; com_pre is resized to 2 x numParts (x = -1, y = 0) when it has another shape */
    void postProcess(cv::Mat& image, Eigen::Matrix<double, 2, Eigen::Dynamic>& com_pre, int interval = 1, int num_threads = 1,
                     cv::Point top_left = cv::Point(0, 0), cv::Point bot_right = cv::Point(-1, -1),
                     double dist_to_pre_weight = 0.001) const;

    std::vector<RNode> nodes;
    std::vector<std::vector<float>> leafData;        // [leaf][numParts]
    std::vector<uint8_t> leafBestMatch;
    int numParts = 0;
    std::vector<int> partMap;
    int partMapType = -1;
   private:
    avb_fitter* device(size_t pixels) const;         // private fitter (on a three-vertex dummy model) with the tree attached
    mutable avb_fitter* fitter_ = nullptr;
    mutable avb_model* dummy_ = nullptr;
    mutable size_t attached_nodes_ = 0;
};

}  // namespace ark
