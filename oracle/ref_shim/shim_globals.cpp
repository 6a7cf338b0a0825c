// TEST INFRASTRUCTURE — storage for the glog shim's flags (see glog/logging.h in this directory).
int FLAGS_logtostderr = 1;
int FLAGS_minloglevel = 0;
int FLAGS_v = 0;
