// Fixture headers in the reference's test format say `#include "ecos.h"`; the names live in ecos_names.h.
#pragma once
#include "ecos_names.h"
