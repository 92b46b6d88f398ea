// TEST INFRASTRUCTURE: src/kind_kernel.hpp includes <loom/pipeline.hpp> but uses nothing from it.
#pragma once
