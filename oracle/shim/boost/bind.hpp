#pragma once
// boost::bind / _1 as used by icp/icp.hpp:390-391, on top of <functional>
#include <functional>
namespace boost { using std::bind; }
using std::placeholders::_1;
