// TEST INFRASTRUCTURE: stands in for the CUDA toolkit header of the same name (tests/emu); the sources use no cooperative-groups API
#pragma once
