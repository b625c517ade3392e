// Empty stand-in. TEST INFRASTRUCTURE ONLY.
#pragma once
#include <vector>
