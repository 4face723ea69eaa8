// SHIM (see CoreMinimal.h in this directory): intentionally empty
#pragma once
#include "CoreMinimal.h"
