// Header name kept from the reference (include/communicate.hpp); the classes live in
// schwz_classes.hpp.
#pragma once
#include "schwz_classes.hpp"
