// capi_util.hpp -- error plumbing of the C ABI: status codes + thread-local message, never exit().
// (The reference prints and calls exit(EXIT_FAILURE) on any CUDA error, helper_cuda.h L566-579.)
#pragma once

#include <string>

#include "../../include/velvet_b200.h"
#include "vt_buffer.hpp"

namespace velvet {

int set_error(int status, const std::string& msg);  // returns status
void clear_error();

}  // namespace velvet

#define VT_API_BEGIN \
    try {            \
        ::velvet::clear_error();
#define VT_API_END                                                          \
    return VELVET_OK;                                                       \
    }                                                                       \
    catch (const ::velvet::Error& e)                                        \
    {                                                                       \
        return ::velvet::set_error(e.status, e.what());                     \
    }                                                                       \
    catch (const std::exception& e)                                         \
    {                                                                       \
        return ::velvet::set_error(VELVET_ERR_STATE, e.what());             \
    }                                                                       \
    catch (...)                                                             \
    {                                                                       \
        return ::velvet::set_error(VELVET_ERR_STATE, "unknown exception");  \
    }
