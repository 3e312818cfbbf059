// Minimal stand-in for the two names fixture headers in the reference's format use.
#pragma once
#include <cstdio>
#include <cstdlib>

inline void mu_assert(const char *message, bool ok)
{
    if (!ok)
    {
        std::printf("%s\n", message);
        std::exit(1);
    }
}
