// gotypes.hpp — the reference's wire structs (internal/cdi/fti/fm/api/*.go, internal/cdi/fti/cm/api/machine.go) as
// type descriptions, so a reply whose JSON types do not fit fails the way json.Unmarshal fails it
// (gojson::decodesInto) instead of being read leniently.
#pragma once
#include "gojson.hpp"

namespace cro {
namespace gotypes {

const gojson::GoType& FMScaleUpResponse();     // fm/api/scale_up.go  ScaleUpResponse
const gojson::GoType& FMGetMachineResponse();  // fm/api/get.go       GetMachineResponse
const gojson::GoType& CMMachineData();         // cm/api/machine.go   MachineData

}  // namespace gotypes
}  // namespace cro
