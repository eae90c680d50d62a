// oracle/shim/boost/filesystem.hpp -- TEST INFRASTRUCTURE: boost::filesystem as std::filesystem (path, exists, operator/,
// directory_iterator) for the reference's AvatarModel.cpp
#pragma once
#include <filesystem>
#include <fstream>   // boost/filesystem.hpp pulls boost/filesystem/fstream.hpp, hence <fstream>; AvatarModel.cpp relies on it
namespace boost {
namespace filesystem {
using namespace std::filesystem;
}
}  // namespace boost
