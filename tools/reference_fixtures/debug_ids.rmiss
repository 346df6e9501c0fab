#version 460
#extension GL_EXT_ray_tracing : require
// Fixture variant of the 3-ray-debug miss stage: white = no hit (instance 15 / primitive 0xfffff is reserved).
layout(location = 0) rayPayloadInEXT vec3 hitValue;

void main()
{
    hitValue = vec3(1.0);
}
