// oracle/shim/boost/filesystem.hpp -- TEST INFRASTRUCTURE: boost::filesystem as std::filesystem (path, exists, operator/,
// directory_iterator) for the reference's AvatarModel.cpp
#pragma once
#include <filesystem>
namespace boost {
namespace filesystem {
using namespace std::filesystem;
}
}  // namespace boost
