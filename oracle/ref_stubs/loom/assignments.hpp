// TEST INFRASTRUCTURE: schedules.hpp includes this header but uses nothing from it.
#pragma once
