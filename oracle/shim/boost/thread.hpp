// oracle/shim/boost/thread.hpp -- TEST INFRASTRUCTURE: boost::thread as std::thread for the reference's AvatarOptimizer.cpp
#pragma once
#include <thread>
namespace boost {
using thread = std::thread;
}
