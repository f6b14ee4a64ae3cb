%% raytracer_gpu_scenes.erl -- synthetic scenes for the GPU tracing functions.
%%
%% "randomly generated scene (not done)" is on the reference's own list (raytracer.erl:35).
%% This module generates the C3/C4 scenes of SURVEY.md section 8(d) as ordinary scene lists
%% (camera first, raytracer.erl:617-619), value for value what
%% eraytracer_b200/scene.py:synthetic_scene/1 produces for the tests and the bench:
%% splitmix64 with the seeds below, u = (x bsr 11) * 2^-53, ten draws per sphere in the order
%% x, y, z, radius, r, g, b, specular_power class, shininess, reflectivity, and every value
%% rounded to binary32 (so a double-precision CPU evaluation and the GPU start from the same
%% numbers).  List order: the three lights, the spheres, the floor plane.
%%
%% NOTE: written without an Erlang/OTP installation (none in the build image): it has never been
%% compiled by erlc.  It IS evaluated by oracle/erlref.py (the Erlang evaluator the reference's own
%% run_tests/0 passes under), and scene(c3) equals the Python generator's scene value for value
%% (tests/test_erl_reference.py::test_erlang_scene_generator_equals_the_python_generator).
-module(raytracer_gpu_scenes).
-export([scene/1, scene/3, splitmix64/2]).

-define(MASK64, 16#FFFFFFFFFFFFFFFF).

%% scene(c3 | c4) -> [Camera | Rest]
scene(c3) -> scene(10000, {{-40.0, 40.0}, {-30.0, 4.0}, {5.0, 85.0}, {0.2, 0.8}}, 16#E7A9C0DE00000003);
scene(c4) -> scene(1000000, {{-200.0, 200.0}, {-150.0, 4.0}, {5.0, 405.0}, {0.2, 1.0}}, 16#E7A9C0DE00000004).

scene(N, {XR, YR, ZR, RR}, Seed) ->
    Camera = {camera, {vector, 0.0, 0.0, -2.0}, {vector, 0.0, 0.0, 0.0}, 90.0, {screen, 4.0, 2.25}},
    Lights = [{point_light, {colour, 1.0, 1.0, 0.5}, {vector, 5.0, -20.0, 0.0}, {colour, 1.0, 1.0, 1.0}},
              {point_light, {colour, 1.0, 0.0, 0.5}, {vector, -30.0, -10.0, 20.0}, {colour, 1.0, 0.0, 0.5}},
              {point_light, {colour, 1.0, 1.0, 1.0}, {vector, 20.0, -40.0, 60.0}, {colour, 1.0, 1.0, 1.0}}],
    Spheres = [sphere(K, XR, YR, ZR, RR, Seed) || K <- lists:seq(0, N - 1)],
    Floor = {plane, {vector, 0.0, -1.0, 0.0}, 5.0,
             {material, {colour, 1.0, 1.0, 1.0}, 1.0, 0.0, f32(0.01)}},
    [Camera | Lights ++ Spheres ++ [Floor]].

%% sphere K uses draws 10*K+1 .. 10*K+10 of the stream
sphere(K, XR, YR, ZR, RR, Seed) ->
    [U0, U1, U2, U3, U4, U5, U6, U7, U8, U9] = [uniform(Seed, 10 * K + J) || J <- lists:seq(1, 10)],
    Power = element(min(trunc(U7 * 4), 3) + 1, {1.0, 4.0, 20.0, 50.0}),
    {sphere, f32(lerp(RR, U3)),
     {vector, f32(lerp(XR, U0)), f32(lerp(YR, U1)), f32(lerp(ZR, U2))},
     {material, {colour, f32(U4), f32(U5), f32(U6)}, Power, f32(U8), f32(U9 * 0.7)}}.

lerp({Lo, Hi}, U) -> Lo + U * (Hi - Lo).

%% nearest binary32 value, as a double
f32(X) -> <<F:32/float>> = <<X:32/float>>, F.

uniform(Seed, K) -> (splitmix64(Seed, K) bsr 11) * math:pow(2.0, -53).

%% output K (1-based) of splitmix64 seeded with Seed
splitmix64(Seed, K) ->
    Z0 = (Seed + K * 16#9E3779B97F4A7C15) band ?MASK64,
    Z1 = ((Z0 bxor (Z0 bsr 30)) * 16#BF58476D1CE4E5B9) band ?MASK64,
    Z2 = ((Z1 bxor (Z1 bsr 27)) * 16#94D049BB133111EB) band ?MASK64,
    Z2 bxor (Z2 bsr 31).
