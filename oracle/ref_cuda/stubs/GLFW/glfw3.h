// stand-in for GLFW: Timer::CurrentTime() (Timer.hpp L89-92) only needs a monotonic clock
#pragma once
#include <chrono>
inline double glfwGetTime()
{
    static const auto t0 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}
