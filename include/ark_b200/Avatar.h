// ark_b200/Avatar.h -- header-compatible stand-in for the reference's include/Avatar.h +
// include/GaussianMixture.h (the members on the fitting path), backed by the avatar_b200 C ABI.
// Same namespace, class names, member names and layouts as include/Avatar.h:64-220, so callers such as
// demo.cpp:135-143,254-268 compile unchanged.  With Eigen/OpenCV installed the real headers are used;
// otherwise ark_b200/mini_eigen.h supplies the handful of types that appear in the signatures.
#pragma once
#if __has_include(<Eigen/Core>)
#include <Eigen/Core>
#include <Eigen/Geometry>
#include <Eigen/Sparse>
#include <Eigen/StdVector>
#else
#include "mini_eigen.h"
#endif
#include <cstdint>
#include <string>
#include <utility>
#include <vector>

struct avb_model;
struct avb_fitter;

namespace ark {
typedef Eigen::Matrix<double, 3, Eigen::Dynamic> CloudType;   // include/Avatar.h:14
typedef Eigen::Matrix<int, 3, Eigen::Dynamic> MeshType;
typedef Eigen::Matrix<double, Eigen::Dynamic, Eigen::Dynamic> MatrixType;

/** include/GaussianMixture.h: only the parsed data lives here; the load() maths runs in the library */
struct GaussianMixture {
    void load(const std::string& path);          // GaussianMixture.cpp:12-58 (text format)
    int numComponents() const { return nComps; }
    int nComps = -1, nDims = 0;
    Eigen::VectorXd weight;
    Eigen::Matrix<double, Eigen::Dynamic, Eigen::Dynamic, Eigen::RowMajor> mean;
    std::vector<Eigen::MatrixXd> cov;
    /** GaussianMixture.cpp:116-132: a random sample (component by weight, then mean + cov_cho * N(0, I)) */
    Eigen::VectorXd sample() const;
};

/** include/Avatar.h:64-151 */
struct AvatarModel {
    explicit AvatarModel(const std::string& model_dir = "", bool limit_one_joint_per_point = false);
    ~AvatarModel();
    AvatarModel(const AvatarModel&) = delete;
    inline int numJoints() const { return (int)parent.rows(); }
    inline int numPoints() const { return (int)(baseCloud.rows() / 3); }
    inline int numShapeKeys() const { return (int)keyClouds.cols(); }
    inline int numFaces() const { return (int)mesh.cols(); }
    inline bool hasMesh() const { return mesh.cols() > 0; }
    inline bool hasPosePrior() const { return posePrior.nComps >= 0; }

    MeshType mesh;
    Eigen::VectorXi parent;
    std::vector<std::vector<std::pair<double, int>>> assignedJoints;
    std::vector<std::vector<std::pair<double, int>>> assignedPoints;
    GaussianMixture posePrior;
    Eigen::VectorXd baseCloud;
    MatrixType keyClouds;
    CloudType initialJointPos;
    Eigen::SparseMatrix<double> jointRegressor;   // kept for source compatibility; not populated by the facade
    bool useJointShapeRegressor = true;
    Eigen::VectorXd jointShapeRegBase;
    Eigen::MatrixXd jointShapeReg;
    Eigen::SparseMatrix<double> weights;          // kept for source compatibility; assignedJoints carries the data
    Eigen::VectorXi assignStarts;                 // declared and never populated in the reference either
    const std::string MODEL_DIR;

    avb_model* handle() const;                    // lazily created avb_model
   private:
    mutable avb_model* handle_ = nullptr;
};

/** include/Avatar.h:155-220 */
class Avatar {
   public:
    explicit Avatar(const AvatarModel& model);
    ~Avatar();
    void update();                                 // Avatar.cpp:22-75 on the GPU
    /** Avatar.cpp:77-126: random shape (N(0,1) per key), pose (a sample of the pose prior) and root position / rotation */
    void randomize(bool randomize_pose = true, bool randomize_shape = true, bool randomize_root_pos_rot = true,
                   uint32_t seed = -1);
    /** Avatar.cpp:141-193: root position / per-joint rotations / shape key 0 that bring the skeleton onto 3 x 24 joint
     *  positions (NaN entries = unknown joint) */
    void alignToJoints(const CloudType& pos);
    Eigen::VectorXd smplParams() const;            // Avatar.cpp:128-137
    const AvatarModel& model;
    CloudType cloud;
    Eigen::VectorXd w;
    Eigen::Vector3d p;
    using Mat3Alloc = Eigen::aligned_allocator<Eigen::Matrix3d>;
    std::vector<Eigen::Matrix3d, Mat3Alloc> r;
    CloudType jointPos;
    Eigen::Matrix<double, 12, Eigen::Dynamic> jointTrans;

    /** x = [p | q xyzw per joint | w] via the optimize() prologue (AvatarOptimizer.cpp:1250-1254) */
    std::vector<double> packParams() const;
    void unpackParams(const std::vector<double>& x);
   private:
    avb_fitter* updater_ = nullptr;
};
}  // namespace ark
