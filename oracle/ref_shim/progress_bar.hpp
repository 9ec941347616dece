/* stand-in for RcppProgress' ProgressBar base class (TEST INFRASTRUCTURE ONLY) */
#ifndef HB_SHIM_PROGRESS_BAR_H
#define HB_SHIM_PROGRESS_BAR_H
class ProgressBar {
 public:
  virtual ~ProgressBar() {}
  virtual void display() = 0;
  virtual void update(float progress) = 0;
  virtual void end_display() = 0;
};
#endif
