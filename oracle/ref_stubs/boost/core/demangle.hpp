// TEST INFRASTRUCTURE — boost::core::demangle is only used for log messages in the MPC sources.
#pragma once
#include <string>
namespace boost { namespace core { inline std::string demangle(const char* n) { return std::string(n); } }}
