// oracle/ref_util.cpp -- TEST INFRASTRUCTURE.  The handful of helpers of the reference's Util.cpp that its AvatarModel.cpp,
// Avatar.cpp and GaussianMixture.cpp link against.  Util.cpp itself cannot be compiled here (it is mostly OpenCV image I/O), so
// these few are RESTATED: the npz array adapters (Util.cpp:250-309: dtype / order dispatch into an Eigen matrix), the root path
// resolver (a stub: the tests always pass a model directory) and the random helpers (deterministic; the tests do not sample).
#include <cstdint>
#include <initializer_list>
#include <random>
#include <string>

#include "Util.h"
#include "UtilCnpy.h"

namespace ark {
namespace util {

namespace {
template <class T, class M>
void fill_from(const cnpy::NpyArray& raw, size_t r, size_t c, M& out) {
    const T* p = raw.data<T>();
    for (size_t i = 0; i < r; ++i)
        for (size_t j = 0; j < c; ++j)
            out(i, j) = (typename M::Scalar)(raw.fortran_order ? p[j * r + i] : p[i * c + j]);
}
}  // namespace

Eigen::Matrix<double, Eigen::Dynamic, Eigen::Dynamic, Eigen::RowMajor> loadFloatMatrix(const cnpy::NpyArray& raw, size_t r, size_t c) {
    _ARK_ASSERT(raw.word_size == 4 || raw.word_size == 8);
    Eigen::Matrix<double, Eigen::Dynamic, Eigen::Dynamic, Eigen::RowMajor> out(r, c);
    if (raw.word_size == 4) fill_from<float>(raw, r, c, out);
    else fill_from<double>(raw, r, c, out);
    return out;
}

Eigen::Matrix<uint32_t, Eigen::Dynamic, Eigen::Dynamic, Eigen::RowMajor> loadUintMatrix(const cnpy::NpyArray& raw, size_t r, size_t c) {
    _ARK_ASSERT(raw.word_size == 4 || raw.word_size == 8);
    Eigen::Matrix<uint32_t, Eigen::Dynamic, Eigen::Dynamic, Eigen::RowMajor> out(r, c);
    if (raw.word_size == 4) fill_from<uint32_t>(raw, r, c, out);
    else fill_from<uint64_t>(raw, r, c, out);
    return out;
}

void assertShape(const cnpy::NpyArray& m, std::initializer_list<size_t> shape) {
    _ARK_ASSERT_EQ(m.shape.size(), shape.size());
    size_t idx = 0;
    for (auto& dim : shape) {
        if (dim != ANY_SHAPE) _ARK_ASSERT_EQ(m.shape[idx], dim);
        ++idx;
    }
}

std::string resolveRootPath(const std::string& root_path) { return root_path; }

}  // namespace util

namespace random_util {
static std::mt19937& gen() { static std::mt19937 g(12345); return g; }
float uniform(float a, float b) { return std::uniform_real_distribution<float>(a, b)(gen()); }
float randn(float m, float v) { return std::normal_distribution<float>(m, v)(gen()); }
float uniform(std::mt19937& rg, float a, float b) { return std::uniform_real_distribution<float>(a, b)(rg); }
float randn(std::mt19937& rg, float m, float v) { return std::normal_distribution<float>(m, v)(rg); }
}  // namespace random_util
}  // namespace ark
