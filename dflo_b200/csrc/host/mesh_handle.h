// the object behind the opaque dflo_mesh handle of include/dflo_host.h
#pragma once
#include "mesh.h"
struct dflo_mesh
{
   dflo::PrimitiveMesh pm;
   dflo::FlatMesh flat;
   dflo_flat_mesh view;
   bool flattened = false;
};
