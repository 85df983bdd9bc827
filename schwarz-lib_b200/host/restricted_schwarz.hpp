// Header name kept from the reference (include/restricted_schwarz.hpp); the classes live in
// schwz_classes.hpp.
#pragma once
#include "schwz_classes.hpp"
