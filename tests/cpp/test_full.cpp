// The reference's driver test_full (src/testmodel.cpp:27-146) written against hpe_classes.hpp:
// same constants, same call sequence (refine_init_pose, pso_evolve, cal_cost, x0 = bestp), frames read from
// Subject1-format files %06d_depth.bin.  Usage: test_full <dir/> <numframes> <num_p> <maxiter> [x0 file] [norefine]
// Prints "frameNNNNNN-cost: c" like the reference plus the pose, one line per frame.
#include <cstdio>
#include <cstdlib>
#include <string>

#include "hpe_classes.hpp"

static const double HGEO_MM[20] = {
    3.110740369831595231e+01, 2.728002169518929776e+01, 2.200017175069322306e+01, 2.199995952496276530e+01,
    6.896344533164219115e+01, 3.080009178775444667e+01, 2.310057694761108138e+01, 2.090010474582363287e+01,
    7.427700757419079025e+01, 3.629982271538526817e+01, 2.420042802617343725e+01, 2.199987377350150197e+01,
    7.744081578614793671e+01, 3.190023249683458673e+01, 2.090000473229609312e+01, 1.979954859510691279e+01,
    8.015924911455944368e+01, 2.310022144042995862e+01, 1.429931098235506148e+01, 1.539998521398316811e+01};
static const double RAD_MM[48] = {
    18.26, 15.18, 16.61, 14.96, 13.53, 12.10, 10.45, 8.80, 12.10, 11.66, 11.8067, 11.9533, 11.11, 10.56, 10.12, 9.68,
    9.46, 9.24, 16.50, 15.0333, 13.5667, 12.10, 11.55, 11.00, 10.78, 10.56, 10.01, 9.46, 16.50, 10.56, 12.54, 14.52,
    10.12, 9.68, 9.46, 9.24, 8.80, 8.36, 12.10, 11.2933, 10.4867, 9.68, 9.46, 9.24, 8.80, 8.36, 7.81, 7.26};

int main(int argc, char **argv)
{
    if (argc < 5) { std::fprintf(stderr, "usage: %s dir/ numframes num_p maxiter [x0 file]\n", argv[0]); return 2; }
    const std::string full = argv[1];
    const int numframes = std::atoi(argv[2]), num_p = std::atoi(argv[3]), maxiter = std::atoi(argv[4]);
    try {
        hpe::vec hgeo(HGEO_MM, HGEO_MM + 20), hrad(RAD_MM, RAD_MM + 48);
        for (double &v : hgeo) v = v / 10.;       // convert to cm, src/testmodel.cpp:49-50
        for (double &v : hrad) v = v / 10.;
        hpe::vec tbnum = {2, 2, 2, 2}, fgnum = {4, 2, 2, 2};
        hpe::vec spc = {-1.86, -1.86, 0, 1.91, 3.84}, hcmc = {150, 107.5, 89.8, 76.5, 59.6};
        hpe::vec x0 = {0, -10, -40, 0, 3, 32, 6, 9, 8, 9, 3, 9, 9, 6, 1, 9, 8, 7, 4, 8, 7, 6, 2, 7, 7, 7};
        if (argc > 5) {                            // optional initial pose (26 doubles, text)
            FILE *f = std::fopen(argv[5], "r");
            for (double &v : x0) if (!f || std::fscanf(f, "%lf", &v) != 1) { std::fprintf(stderr, "bad x0 file\n"); return 2; }
            std::fclose(f);
        }
        hpe::handmodel hand(hgeo, spc, tbnum, fgnum, hcmc, hrad);

        double focal = 241.42;
        int imgH = 320, imgW = 240;
        hpe::observedmodel observation;
        observation.init_observation(full, "000000_depth.bin", true, imgW, imgH, focal, true);
        hpe::costfunc optfunc(&hand, &observation);

        hpe::vec ub(26), lb(26), std_(26);
        const double tu[4] = {15, 90, 110, 90}, tl[4] = {-15, 0, 0, 0};
        for (int i = 0; i < 26; ++i) {
            ub[i] = i < 3 ? 180 : i < 6 ? 100 : tu[(i - 6) % 4];
            lb[i] = i < 3 ? -180 : i < 6 ? -100 : tl[(i - 6) % 4];
            std_[i] = i < 3 ? 9.0 : i < 6 ? 7.0 : 9.0;
        }
        hpe::PSO optimiser;
        optimiser.set_pso_params(ub, lb, std_, 0.7298, 1.49618, 1.49618, maxiter, 1e-8, 1e-8);

        for (int frame = 0; frame < numframes; frame++) {
            char name[64];
            std::snprintf(name, sizeof name, "%06d_depth.bin", frame);
            observation.next_frame(name);
            if (argc <= 6) optimiser.refine_init_pose(x0, optfunc);      // src/testmodel.cpp:126
            hpe::vec bestp;
            optimiser.pso_evolve(optfunc, x0, num_p, bestp);
            double c = optfunc.cal_cost(bestp);
            std::printf("frame%06d-cost: %.17g pose:", frame, c);
            for (double v : bestp) std::printf(" %.17g", v);
            std::printf("\n");
            x0 = bestp;
        }
        // the getters of observedmodel / handmodel
        hpe::mat spheres;
        hand.build_hand_model(x0, spheres);
        std::printf("scale %.17g joint4 %.17g %.17g %.17g sphere47 %.17g\n", observation.get_img_scale(),
                    hand.hand_joints(4, 0), hand.hand_joints(4, 1), hand.hand_joints(4, 2), spheres(47, 2));
    } catch (const hpe::error &e) {
        std::fprintf(stderr, "hpe error %d: %s\n", e.code, e.what());
        return 1;
    }
    return 0;
}
