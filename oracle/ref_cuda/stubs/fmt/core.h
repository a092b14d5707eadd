// stand-in for {fmt}: the reference only prints diagnostics with it
#pragma once
#include <string>
namespace fmt {
template <class... A> inline void print(A&&...) {}
template <class... A> inline std::string format(A&&...) { return std::string(); }
}
